#!/usr/bin/env python3
"""Developer check on the GPU box: render time of the three pipelines on several workloads (same images)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)
import lisa_b200.frontend as fe, lisa_b200.rt as rt

def run(sc, pipe, spp, count=1, reps=2):
    if pipe == "path": os.environ.pop("LISA_PIPELINE", None)
    else: os.environ["LISA_PIPELINE"] = pipe
    R = rt.Renderer.from_scene(sc)
    R.render_subframes(100, 1, 1); R.reset()
    best = 1e30
    for r in range(reps):
        R.reset(); R.render_subframes(0, count, spp)
        best = min(best, R.stats()["last_render_ms"])
    img = R.read_accum(); R.close()
    os.environ.pop("LISA_PIPELINE", None)
    return best, img

cases = []
c1 = fe.parse_scene("scenes/cornell_c1.rto")
cases.append(("C1 512x512 64 spp", c1, 64, 1))
tiny = dict(c1); tiny["width"] = tiny["height"] = 128
cases.append(("Cornell 128x128 64 spp", tiny, 64, 1))
cases.append(("Cornell 128x128 16 spp x 8 subframes", tiny, 16, 8))
if os.path.exists("out/knot.obj"):
    cases.append(("C3 1920x1080 16 spp", fe.parse_scene("scenes/c3_knot.rto"), 16, 1))
for name, sc, spp, count in cases:
    res = {}
    for pipe in ("pool", "path", "wavefront"):
        ms, img = run(sc, pipe, spp, count)
        res[pipe] = (ms, img)
    same = all(np.array_equal(res["pool"][1], res[p][1]) for p in ("path", "wavefront"))
    print("%-40s pool %.2f ms  path %.2f ms  wavefront %.2f ms  identical=%s" % (name, res["pool"][0], res["path"][0], res["wavefront"][0], same), flush=True)
