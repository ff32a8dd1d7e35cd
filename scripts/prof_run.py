#!/usr/bin/env python3
"""Short run of the hot path for ncu: README Cornell box at 2000x2000, a few spp (same kernels, same grid
sizes and the same per-iteration work as bench.py's steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lisa_b200.frontend as fe
import lisa_b200.rt as rt
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
sc = fe.parse_scene("scenes/cornell_c2.rto")
sc["width"] = sc["height"] = w
R = rt.Renderer.from_scene(sc, bvh_kind=int(os.environ.get("PROF_BVH", "0")))
if os.environ.get('PROF_WARM', '1') == '1' and 'NCU' not in os.environ:
    R.render_subframes(99, 1, 2); R.reset()
R.render_subframes(0, 1, spp)
st = R.stats()
print("render %.1f ms, %.2f Msamples/s, %.1f Mrays/s, launches %d" % (st["last_render_ms"], st["last_samples"] / st["last_render_ms"] / 1e3,
      (st["last_radiance_rays"] + st["last_shadow_rays"]) / st["last_render_ms"] / 1e3, st["last_kernel_launches"]))
