#!/usr/bin/env python3
"""Developer check on the GPU box: two library variants on the 1M-triangle soup — how many pixels differ (ties in t between
overlapping triangles resolve by test order) and closest hits against the oracle for each.  usage: leaf_check.py LIB_A LIB_B"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import lisa_b200.rt as rt
    from oracle.scene_py import pack_material
    from scripts.scale_bench import check_vs_oracle
    T = 1_000_000
    rng = np.random.default_rng(0x5EED)
    edge = 0.5 * T ** (-1.0 / 3.0)
    c = rng.random((T, 1, 3), dtype=np.float32)
    v = (c + (rng.random((T, 3, 3), dtype=np.float32) - 0.5) * np.float32(2 * edge)).reshape(-1, 3)
    e1 = v[1::3] - v[0::3]; e2 = v[2::3] - v[0::3]
    fn = np.cross(e1, e2); fn /= (np.linalg.norm(fn, axis=1, keepdims=True) + 1e-30)
    n = np.repeat(fn.astype(np.float32), 3, axis=0)
    q = np.float32([[-0.5, 1.6, -0.5], [1.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, 1.5]])
    v = np.concatenate([v, q]); n = np.concatenate([n, np.tile(np.float32([[0, -1, 0]]), (6, 1))])
    m = np.concatenate([np.zeros(T, np.int32), np.ones(2, np.int32)])
    mats = [dict(emit=False, alpha=1.0, diffuse=(0.7, 0.7, 0.7), roughness=1.0), dict(emit=True, alpha=1.0, emission=(1, 1, 1))]
    R = rt.Renderer(v, n, m, mats, 512, 512, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, 1, 7)
    R.render_subframes(0, 1, 1)
    acc = R.read_accum(); st = R.stats()
    mp = b"".join(pack_material(roughness=x.get("roughness", 0), alpha=x["alpha"], diffuse=x.get("diffuse", (0, 0, 0)), emit=x["emit"], emission=x.get("emission", (0, 0, 0))) for x in mats)
    bad, nn = check_vs_oracle(R, v, n, m, mp, 2000)
    np.save(sys.argv[2], acc)
    print(json.dumps({"bad": bad, "n": nn, "nodes": st["bvh_nodes"], "nodes_visited": st["last_nodes_visited"], "tris_tested": st["last_triangles_tested"], "rays": st["last_radiance_rays"] + st["last_shadow_rays"]}))
    sys.exit(0)
import numpy as np
outs = []
for i, lib in enumerate(sys.argv[1:3]):
    env = dict(os.environ)
    if lib != "product": env["LISA_RT_LIB"] = os.path.join(ROOT, "lisa_b200", "variants", "liblisa_rt_%s.so" % lib)
    f = "/tmp/_leaf_%d.npy" % i
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", f], env=env, capture_output=True, text=True, cwd=ROOT)
    print(lib, p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-500:], flush=True)
    outs.append(np.load(f))
d = np.abs(outs[0][..., :3] - outs[1][..., :3]).max(axis=2)
print("pixels differing: %d of %d (%.4f %%), mean ratio %.6f" % ((d > 0).sum(), d.size, 100.0 * (d > 0).mean(), outs[1][..., :3].mean() / outs[0][..., :3].mean()))
