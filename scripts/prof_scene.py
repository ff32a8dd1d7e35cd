#!/usr/bin/env python3
"""One render of a larger BASELINE scene for ncu (the k_pool launch is what gets captured):
  prof_scene.py knot [spp]      C3: 871,218 triangles, 1920x1080, 12 bounces
  prof_scene.py soup T [spp]    C4: T-triangle soup, 1024x1024, 7 bounces"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)
import lisa_b200.frontend as fe, lisa_b200.rt as rt
os.environ.setdefault("LISA_PIPELINE", "pool")
if sys.argv[1] == "knot":
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    sys.path.insert(0, os.path.join(ROOT, "assets"))
    import gen_knot
    sc = fe.parse_scene("scenes/cornell_c1.rto")
    kv, kn = gen_knot.soup_arrays()
    T0 = len(sc["mat_indices"])
    keep = np.isin(np.arange(T0), np.r_[0:16, T0 - 2:T0])
    v = np.concatenate([sc["vertices"].reshape(-1, 3, 3)[keep].reshape(-1, 3), kv])
    n = np.concatenate([sc["normals"].reshape(-1, 3, 3)[keep].reshape(-1, 3), kn])
    m = np.concatenate([sc["mat_indices"][keep], np.full(len(kv) // 3, 1, np.int32)])
    R = rt.Renderer(v, n, m, sc["materials_packed"], 1920, 1080, sc["camera"]["eye"], sc["camera"]["look_at"], sc["camera"]["fov"], spp, 12)
else:
    from oracle.make_golden_gpu import soup   # the C4 generator (test infrastructure; scripts may use it)
    T = int(float(sys.argv[2])); spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    v, n = soup(T)
    q = np.float32([[-0.5, 1.6, -0.5], [1.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, 1.5]])
    v = np.concatenate([v, q]); n = np.concatenate([n, np.tile(np.float32([[0, -1, 0]]), (6, 1))])
    m = np.concatenate([np.zeros(T, np.int32), np.ones(2, np.int32)])
    mats = [dict(emit=False, alpha=1.0, diffuse=(0.7, 0.7, 0.7), roughness=1.0), dict(emit=True, alpha=1.0, emission=(1, 1, 1))]
    R = rt.Renderer(v, n, m, mats, 1024, 1024, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, spp, 7)
R.render_subframes(0, 1, spp)
st = R.stats()
rays = st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]
print("render %.1f ms, %.2f Msamples/s, %.1f Mrays/s traversed, %.2f nodes + %.2f triangles per ray" % (
    st["last_render_ms"], st["last_samples"] / st["last_render_ms"] / 1e3, rays / st["last_render_ms"] / 1e3,
    st["last_nodes_visited"] / max(rays, 1), st["last_triangles_tested"] / max(rays, 1)))
