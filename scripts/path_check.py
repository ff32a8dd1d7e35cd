#!/usr/bin/env python3
"""Developer check on the GPU box: the persistent pipeline against the wavefront pipeline (bit-identical images and
counters), then timings of both on BASELINE configs[1] (Cornell 2000x2000, 7 bounces) in steps of 50 spp."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lisa_b200.frontend as fe
import lisa_b200.rt as rt

def timing(sc, flags, spp=50, reps=3):
    R = rt.Renderer.from_scene(sc, flags=flags)
    R.render_subframes(0, 1, 4)
    ms = []
    for i in range(reps):
        R.render_subframes(1 + i, 1, spp)
        ms.append(R.stats()["last_render_ms"])
    st = R.stats()
    R.close()
    n = sc["width"] * sc["height"] * spp
    return min(ms), n / min(ms) / 1e3, st

if __name__ == "__main__":
    sc = fe.parse_scene("scenes/cornell_c1.rto")
    if "--skip-parity" not in sys.argv:
        for w, spp in ((64, 16), (256, 8)):
            s2 = dict(sc); s2["width"] = s2["height"] = w
            out = []
            for fl in (0, rt.FLAG_WAVEFRONT):
                R = rt.Renderer.from_scene(s2, flags=fl)
                R.render_subframes(0, 2, spp)
                out.append((R.read_accum(), R.stats()))
                R.close()
            same = np.array_equal(out[0][0], out[1][0])
            d = np.abs(out[0][0] - out[1][0]).max()
            print("parity %dx%d %d spp: identical=%s maxdiff=%g  rays %d/%d shadow %d/%d culled %d/%d" % (
                w, w, spp, same, d, out[0][1]["last_radiance_rays"], out[1][1]["last_radiance_rays"],
                out[0][1]["last_shadow_rays"], out[1][1]["last_shadow_rays"], out[0][1]["last_shadow_culled"], out[1][1]["last_shadow_culled"]), flush=True)
    c2 = fe.parse_scene("scenes/cornell_c2.rto")
    for name, fl in (("path", 0), ("wavefront", rt.FLAG_WAVEFRONT)):
        if name == "wavefront" and "--no-wavefront" in sys.argv:
            continue
        ms, msps, st = timing(c2, fl)
        print("%s: %.1f ms per 50 spp  %.1f Msamples/s  launches %d  nodes/tris per traversed ray %.2f %.2f" % (
            name, ms, msps, st["last_kernel_launches"],
            st["last_nodes_visited"] / max(1, st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]),
            st["last_triangles_tested"] / max(1, st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"])), flush=True)
