#!/usr/bin/env python3
"""Developer check on the GPU box: k_cta (LISA_PIPELINE=cta) against k_pool — bit-identical images and counters —
then timings of both on BASELINE configs[1]."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lisa_b200.frontend as fe
import lisa_b200.rt as rt

def render(sc, pipe, first, count, spp, **kw):
    os.environ["LISA_PIPELINE"] = pipe
    R = rt.Renderer.from_scene(sc, **kw)
    R.render_subframes(first, count, spp)
    out = (R.read_accum(), R.stats())
    R.close()
    return out

sc = fe.parse_scene("scenes/cornell_c1.rto")
if "--skip-parity" not in sys.argv:
    for w, spp, kw in ((8, 2, {}), (64, 16, {}), (256, 8, {}), (128, 8, dict(bvh_kind=1)), (128, 8, dict(shadow_mode=1)), (128, 4, dict(flags=rt.FLAG_NO_CULL)), (40, 3, {}), (700, 3, {})):
        s2 = dict(sc); s2["width"] = w; s2["height"] = w if w != 40 else 33
        a, sa = render(s2, "cta", 0, 2, spp, **kw)
        b, sb = render(s2, "pool", 0, 2, spp, **kw)
        keys = ("last_radiance_rays", "last_shadow_rays", "last_shadow_culled", "last_shadow_jobs", "null_directions", "last_nodes_visited", "last_triangles_tested", "last_samples")
        print("parity %dx%d %d spp %s: identical=%s maxdiff=%g counters %s" % (w, s2["height"], spp, kw, np.array_equal(a, b), np.abs(a - b).max(),
              "equal" if all(sa[k] == sb[k] for k in keys) else {k: (sa[k], sb[k]) for k in keys if sa[k] != sb[k]}), flush=True)
c2 = fe.parse_scene("scenes/cornell_c2.rto")
for pipe in ("cta", "pool", "cta", "pool"):
    os.environ["LISA_PIPELINE"] = pipe
    R = rt.Renderer.from_scene(c2)
    R.render_subframes(0, 1, 4)
    ms = []
    for i in range(3):
        R.render_subframes(1 + i, 1, 50)
        ms.append(R.stats()["last_render_ms"])
    R.close()
    print("%s: %s ms -> %.1f Msamples/s" % (pipe, ["%.1f" % m for m in ms], 2000 * 2000 * 50 / min(ms) / 1e3), flush=True)
