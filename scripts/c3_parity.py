#!/usr/bin/env python3
"""C3 (871k-triangle glass knot in the Cornell box, 12 bounces) against the UNMODIFIED reference OptiX renderer on the
same box: per-pixel agreement at equal spp with the reference's seeding, image means, and both render times.
Run on the GPU box:  python assets/gen_knot.py out/knot.obj && python scripts/c3_parity.py"""
import json, os, re, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)
import lisa_b200.frontend as fe, lisa_b200.rt as rt

def variant(w, h, spp):
    txt = open("scenes/c3_knot.rto").read()
    if "--no-ceiling" in sys.argv:  # Q2 probe: nothing behind the light, so first-found == closest for rays that reach it
        txt = re.sub(r"mesh \{\s*obj_file = assets/objs/cornell_box/top.obj\s*material = white\s*\}", "", txt)
    txt = re.sub(r"num_samples = \d+", "num_samples = %d" % spp, txt)
    txt = re.sub(r"width = \d+", "width = %d" % w, txt); txt = re.sub(r"height = \d+", "height = %d" % h, txt)
    p = "out/c3_%dx%d_%d.rto" % (w, h, spp); open(p, "w").write(txt); return p

def optix(scene, accum):
    r = subprocess.run(["oracle/_ref/lisa_optix_ref", "-s", scene, "--warmup", "--accum", accum], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=3000)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode or not line: raise RuntimeError(r.stderr[-500:])
    return json.loads(line[-1])

out = {}
for (w, h, spp) in ((480, 270, 1), (480, 270, 16)) + (((1920, 1080, 16),) if '--full' in sys.argv else ()):
    sc_path = variant(w, h, spp)
    ref = optix(sc_path, "out/c3_ref.f32")
    ref_img = np.fromfile("out/c3_ref.f32", dtype=np.float32).reshape(h, w, 4)[..., :3]
    sc = fe.parse_scene(sc_path)
    R = rt.Renderer.from_scene(sc)
    R.render_subframes(100, 1, 1); R.reset()
    R.render()
    st = R.stats(); img = R.read_accum()[..., :3]
    d = np.abs(img - ref_img).max(axis=2)
    if w == 480:
        np.savez_compressed('gpurun_out/c3_%d%s.npz' % (spp, '_nc' if '--no-ceiling' in sys.argv else ''), ours=img, ref=ref_img)
    out["%dx%d_%dspp" % (w, h, spp)] = dict(
        pixels_within_1e_4=float((d < 1e-4).mean()), mean_ours=[float(x) for x in img.reshape(-1, 3).mean(0)], mean_ref=ref["mean_rgb"],
        mean_ratio=[float(a / b) for a, b in zip(img.reshape(-1, 3).mean(0), ref["mean_rgb"])],
        ours_render_ms=round(st["last_render_ms"], 2), ours_msamples_per_s=round(st["last_samples"] / st["last_render_ms"] / 1e3, 2),
        ours_bvh_build_ms=round(st["bvh_build_ms"], 2), ref_render_ms=ref["render_ms"], ref_msamples_per_s=ref["msamples_per_s"],
        ref_setup_ms=ref["setup_ms"], speedup=round(ref["render_ms"] / st["last_render_ms"], 2))
    R.close()
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/c3_parity%s.json" % ("_noceiling" if "--no-ceiling" in sys.argv else ""), "w"), indent=1)
