#!/usr/bin/env python3
"""BVH build time of the BASELINE C4 soups (and optionally the C3 knot), warm: the stage breakdown lisa_create prints under
LISA_DEBUG_TIMING (stderr) plus one JSON line per scene with the build time from CUDA events, the node count and the
tree-quality counters of a short render.   python scripts/build_bench.py 1000000 10000000 [100000000] [--lbvh]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)
os.environ["LISA_DEBUG_TIMING"] = "1"
import lisa_b200.rt as rt
from oracle.make_golden_gpu import soup   # the C4 generator (test infrastructure: scripts may use it)

mats = [dict(emit=False, alpha=1.0, diffuse=(0.7, 0.7, 0.7), roughness=1.0), dict(emit=True, alpha=1.0, emission=(1, 1, 1))]
q = np.float32([[-0.5, 1.6, -0.5], [1.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, 1.5]])
flags = rt.FLAG_LBVH if "--lbvh" in sys.argv else 0
for a in sys.argv[1:]:
    if a.startswith("--"):
        continue
    T = int(float(a))
    v, n = soup(T)
    v = np.concatenate([v, q]); n = np.concatenate([n, np.tile(np.float32([[0, -1, 0]]), (6, 1))])
    m = np.concatenate([np.zeros(T, np.int32), np.ones(2, np.int32)])
    builds = []
    for k in range(3):
        sys.stderr.write("---- %d triangles, build %d\n" % (T, k)); sys.stderr.flush()
        R = rt.Renderer(v, n, m, mats, 1024, 1024, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, 1, 7, flags=flags)
        st = R.stats()
        builds.append(round(st["bvh_build_ms"], 3))
        if k < 2:
            R.close()
    R.render_subframes(0, 1, 2)
    st = R.stats()
    rays = st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]
    print(json.dumps(dict(triangles=T, bvh_build_ms=builds, upload_ms=round(st["upload_ms"], 2), bvh_nodes=st["bvh_nodes"],
                          nodes_per_ray=round(st["last_nodes_visited"] / max(rays, 1), 2), tris_per_ray=round(st["last_triangles_tested"] / max(rays, 1), 2),
                          render_ms_2spp=round(st["last_render_ms"], 2), msamples_per_s=round(st["last_samples"] / st["last_render_ms"] / 1e3, 2))), flush=True)
    R.close()
    del v, n, m
