#!/usr/bin/env python3
"""Golden fixtures and reference timings for the larger BASELINE configs — run ON THE GPU BOX.  TEST INFRASTRUCTURE.

  gpurun -- 'python oracle/make_golden_gpu.py [--timings]'

Runs the UNMODIFIED reference OptiX renderer (oracle/_ref/lisa_optix_ref, built from /root/reference by oracle/Makefile)
on each scene, reduces its float4 accumulators to small fixtures (mean, block means, a 128x128 crop) under
gpurun_out/golden/ — copy them to tests/golden/ — and renders the same scene with the CUDA path to report the agreement
the tests in tests/test_gpu_configs.py then assert (gpurun_out/golden/report.json):

  optix_c2_4        BASELINE configs[1] resolution: README Cornell box 2000x2000, 4 spp, 7 bounces
  optix_c3_1 / _16  BASELINE configs[2]: 871,200-triangle glass knot in the Cornell box, 12 bounces, 480x270, 1 / 16 spp
  optix_c3q2_16     the same with the ORIGINAL ceiling (scenes/c3_knot_q2.rto): the Q2 sensitivity fixture
  optix_c4_1m_4     BASELINE configs[3]: 1M-triangle soup, 256x256, 4 spp, 7 bounces (soup via --soup-bin)
--timings adds the reference's setup (upload + optixAccelBuild + pipeline) and render times for the 1M / 10M soups at
1024x1024 and the 4K Cornell config beside ours (gpurun_out/golden/ref_timings.json).
"""
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "lisa_optix_ref")

SOUP_RTO = """/* BASELINE configs[3]: procedural triangle soup (appended by the caller) + one emissive quad above the unit cube */
material grey {
  color = (0.7, 0.7, 0.7, 1)
  roughness = 1
}
material light {
  emit = true
  color = (1, 1, 1, 1)
}
mesh {
  obj_file = %(light)s
  material = light
}
camera {
  position = (0.5, 0.6, 3.2)
  look_at = (0.5, 0.45, 0.5)
  fov = 35
}
num_samples = %(spp)d
num_bounces = 7
width = %(w)d
height = %(h)d
output_image = out/soup.ppm
"""
LIGHT_OBJ = """# emissive quad above the unit cube, normal -y
v -0.5 1.6 -0.5
v 1.5 1.6 -0.5
v 1.5 1.6 1.5
v -0.5 1.6 1.5
vn 0 -1 0
f 1//1 2//1 3//1
f 1//1 3//1 4//1
"""


def soup(T, seed=0x5EED):
    """BASELINE C4 soup: centres uniform in the unit cube, edge ~ 0.5 T^(-1/3), random orientation, flat normals."""
    rng = np.random.default_rng(seed)
    edge = 0.5 * T ** (-1.0 / 3.0)
    c = rng.random((T, 1, 3), dtype=np.float32)
    v = (c + (rng.random((T, 3, 3), dtype=np.float32) - 0.5) * np.float32(2 * edge)).reshape(-1, 3)
    e1, e2 = v[1::3] - v[0::3], v[2::3] - v[0::3]
    fn = np.cross(e1, e2)
    fn /= (np.linalg.norm(fn, axis=1, keepdims=True) + 1e-30)
    return v.astype(np.float32), np.repeat(fn.astype(np.float32), 3, axis=0)


def write_soup_scene(T, w, h, spp, tag):
    os.makedirs("out", exist_ok=True)
    light = "out/soup_light.obj"
    open(light, "w").write(LIGHT_OBJ)
    rto = "out/soup_%s.rto" % tag
    open(rto, "w").write(SOUP_RTO % dict(light=light, spp=spp, w=w, h=h))
    v, n = soup(T)
    binp = "out/soup_%s.bin" % tag
    with open(binp, "wb") as f:
        f.write(np.uint64(T).tobytes())
        f.write(v.tobytes())
        f.write(n.tobytes())
    return rto, binp, v, n


def variant(src, w, h, spp, tag):
    txt = open(src).read()
    txt = re.sub(r"num_samples = \d+", "num_samples = %d" % spp, txt)
    txt = re.sub(r"width = \d+", "width = %d" % w, txt)
    txt = re.sub(r"height = \d+", "height = %d" % h, txt)
    p = "out/%s.rto" % tag
    open(p, "w").write(txt)
    return p


def optix(scene, accum=None, extra=()):
    cmd = [REF, "-s", scene, "--warmup"] + (["--accum", accum] if accum else []) + list(extra)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=3000)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode or not line:
        raise RuntimeError("optix_ref failed on %s: %s" % (scene, r.stderr[-600:]))
    return json.loads(line[-1])


def reduce_image(a, block):
    h, w, _ = a.shape
    out = {"mean_rgb": a.reshape(-1, 3).mean(axis=0, dtype=np.float64), "block": np.array(block)}
    hb, wb = h // block * block, w // block * block
    out["block_mean"] = a[:hb, :wb].reshape(hb // block, block, wb // block, block, 3).mean(axis=(1, 3), dtype=np.float64).astype(np.float32)
    y0, x0 = max(0, h // 2 - 64), max(0, w // 2 - 64)
    out["crop"] = a[y0:y0 + 128, x0:x0 + 128].copy()
    out["crop_origin"] = np.array([y0, x0])
    return out


def agreement(ours, ref, g):
    d = np.abs(ours - ref).max(axis=2)
    y0, x0 = g["crop_origin"]
    b = int(g["block"])
    hb, wb = ours.shape[0] // b * b, ours.shape[1] // b * b
    bm = ours[:hb, :wb].reshape(hb // b, b, wb // b, b, 3).mean(axis=(1, 3))
    return {"pixels_within_1e-4": float((d < 1e-4).mean()),
            "crop_pixels_within_1e-4": float((d[y0:y0 + 128, x0:x0 + 128] < 1e-4).mean()),
            "mean_ratio": [float(x) for x in ours.reshape(-1, 3).mean(0) / np.maximum(ref.reshape(-1, 3).mean(0), 1e-12)],
            "block_mean_max_abs_diff": float(np.abs(bm - g["block_mean"]).max())}


def main():
    import lisa_b200.frontend as fe
    import lisa_b200.rt as rt
    os.makedirs(OUT, exist_ok=True)
    os.makedirs("out", exist_ok=True)
    report = {}

    def ours_render(sc, spp, **kw):
        R = rt.Renderer.from_scene(sc, **kw)
        R.render_subframes(0, 1, spp)
        img, st = R.read_accum()[..., :3].copy(), R.stats()
        R.close()
        return img, st

    def case(name, scene, w, h, spp, block, sc_fn, extra=(), policies=(0,)):
        accum = "out/%s.f32" % name
        ref = optix(scene, accum, extra)
        a = np.fromfile(accum, dtype=np.float32).reshape(h, w, 4)[..., :3]
        os.remove(accum)
        g = reduce_image(a, block)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
        sc = sc_fn()
        rep = {"ref": {k: ref[k] for k in ("width", "height", "spp", "bounces", "triangles", "setup_ms", "render_ms", "msamples_per_s", "mean_rgb")}}
        for pol in policies:
            for pipe in ("path", "pool"):
                os.environ["LISA_PIPELINE"] = pipe
                img, st = ours_render(sc, spp, shadow_mode=pol)
                key = "%s_%s" % ("closest" if pol == 0 else "first_found", pipe)
                rep[key] = dict(agreement(img, a, g), render_ms=round(st["last_render_ms"], 3), bvh_build_ms=round(st["bvh_build_ms"], 3),
                                nodes_per_ray=round(st["last_nodes_visited"] / max(1, st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]), 2))
            os.environ.pop("LISA_PIPELINE", None)
        report[name] = rep
        print(name, json.dumps(rep), flush=True)
        json.dump(report, open(os.path.join(OUT, "report.json"), "w"), indent=1)

    def parsed(path):
        return lambda: fe.parse_scene(path)

    # ---- C2 resolution
    p = variant("scenes/cornell_c2.rto", 2000, 2000, 4, "g_c2_4")
    case("optix_c2_4", p, 2000, 2000, 4, 8, parsed(p))
    # ---- C3 (the knot OBJ is generated, 75 MB: never committed)
    if not os.path.exists("out/knot.obj"):
        subprocess.check_call([sys.executable, "assets/gen_knot.py", "out/knot.obj"])
    for spp in (1, 16):
        p = variant("scenes/c3_knot.rto", 480, 270, spp, "g_c3_%d" % spp)
        case("optix_c3_%d" % spp, p, 480, 270, spp, 10, parsed(p))
    p = variant("scenes/c3_knot_q2.rto", 480, 270, 16, "g_c3q2_16")
    case("optix_c3q2_16", p, 480, 270, 16, 10, parsed(p), policies=(0, 1))
    # ---- C4, 1M soup
    rto, binp, v, n = write_soup_scene(1_000_000, 256, 256, 4, "1m_256")

    def soup_scene(rto=rto, v=v, n=n):
        sc = fe.parse_scene(rto)
        T = v.shape[0] // 3
        sc["vertices"] = np.concatenate([v, sc["vertices"]])
        sc["normals"] = np.concatenate([n, sc["normals"]])
        sc["mat_indices"] = np.concatenate([np.zeros(T, np.int32), sc["mat_indices"]])
        return sc
    case("optix_c4_1m_4", rto, 256, 256, 4, 8, soup_scene, extra=("--soup-bin", binp, "--soup-material", "0"))

    if "--timings" in sys.argv:
        tim = {}
        for T, tag, spp in ((1_000_000, "1m", 16), (10_000_000, "10m", 4)):
            rto, binp, v, n = write_soup_scene(T, 1024, 1024, spp, tag + "_1024")
            t0 = time.perf_counter()
            ref = optix(rto, None, ("--soup-bin", binp, "--soup-material", "0"))
            sc = soup_scene(rto, v, n)
            R = rt.Renderer.from_scene(sc)
            R.render_subframes(99, 1, 1); R.reset()
            R.render_subframes(0, 1, spp)
            st = R.stats()
            R.close()
            R = rt.Renderer.from_scene(sc)   # warm build (allocator cache, modules loaded)
            st2 = R.stats()
            R.close()
            tim["c4_soup_" + tag] = {"triangles": T, "width": 1024, "height": 1024, "spp": spp,
                                     "ref_setup_ms_upload_accelbuild_pipeline": ref["setup_ms"], "ref_render_ms": ref["render_ms"], "ref_msamples_per_s": ref["msamples_per_s"],
                                     "ours_upload_ms": round(st2["upload_ms"], 2), "ours_bvh_build_ms_warm": round(st2["bvh_build_ms"], 2),
                                     "ours_render_ms": round(st["last_render_ms"], 2), "ours_msamples_per_s": round(st["last_samples"] / st["last_render_ms"] / 1e3, 2),
                                     "mean_ratio_ours_over_ref": None, "wall_s": round(time.perf_counter() - t0, 1)}
            print(json.dumps(tim), flush=True)
            del v, n, sc
        # C5: 3840x2160 Cornell; bounded sample of the 4096 spp (64 spp), both arms
        p = variant("scenes/cornell_4k.rto", 3840, 2160, 64, "g_c5_64")
        ref = optix(p)
        sc = fe.parse_scene(p)
        R = rt.Renderer.from_scene(sc)
        R.render_subframes(99, 1, 1); R.reset()
        R.render_subframes(0, 1, 64)
        st = R.stats()
        R.close()
        tim["c5_cornell_4k_64spp"] = {"ref_render_ms": ref["render_ms"], "ref_msamples_per_s": ref["msamples_per_s"], "ref_setup_ms": ref["setup_ms"],
                                      "ours_render_ms": round(st["last_render_ms"], 2), "ours_msamples_per_s": round(st["last_samples"] / st["last_render_ms"] / 1e3, 2),
                                      "note": "64 of the config's 4096 spp (both arms scale linearly in spp): full config = x64"}
        p = variant("scenes/c3_knot.rto", 1920, 1080, 16, "g_c3_full16")
        ref = optix(p)
        sc = fe.parse_scene(p)
        R = rt.Renderer.from_scene(sc)
        R.render_subframes(99, 1, 1); R.reset()
        R.render_subframes(0, 1, 16)
        st = R.stats()
        R.close()
        tim["c3_knot_1080p_16spp"] = {"ref_render_ms": ref["render_ms"], "ref_msamples_per_s": ref["msamples_per_s"], "ref_setup_ms": ref["setup_ms"],
                                      "ours_render_ms": round(st["last_render_ms"], 2), "ours_msamples_per_s": round(st["last_samples"] / st["last_render_ms"] / 1e3, 2),
                                      "ours_bvh_build_ms": round(st["bvh_build_ms"], 2)}
        json.dump(tim, open(os.path.join(OUT, "ref_timings.json"), "w"), indent=1)
        print(json.dumps(tim, indent=1))


if __name__ == "__main__":
    main()
