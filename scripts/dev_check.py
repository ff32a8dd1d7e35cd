#!/usr/bin/env python3
"""Developer check run on the GPU box: BVH queries vs the oracle, image parity vs the reference OptiX
fixtures, and quick timings.  Not part of the product; imports oracle/ as the checker."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import scene_py, binding
import lisa_b200.rt as rt

G = "tests/golden"

def load(scene):
    return scene_py.parse_scene(scene)

def check_bvh(sc, bvh_kind, n=200000):
    R = rt.Renderer.from_scene(sc, bvh_kind=bvh_kind)
    print("stats", {k: v for k, v in R.stats().items() if k in ("num_triangles", "num_emitter_triangles", "bvh_nodes", "bvh_bytes", "bvh_build_ms", "upload_ms")})
    rng = np.random.default_rng(1)
    lo, hi = sc["vertices"].min(0), sc["vertices"].max(0)
    org = rng.uniform(lo - 0.05, hi + 0.05, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.uniform(0.3, 2.0, size=(n, 1)).astype(np.float32)
    prim, t = R.trace_closest(org, d)
    S = binding.Scene(sc["vertices"], sc["normals"], sc["mat_indices"], sc["materials_packed"])
    m = 20000
    bad = 0
    for i in range(m):
        p, tt = S.closest_hit(org[i], d[i])
        if p != prim[i]:
            # allow ties: same t
            if p >= 0 and prim[i] >= 0 and abs(tt - t[i]) <= 1e-5 * max(1, abs(tt)):
                continue
            bad += 1
            if bad < 5: print("mismatch", i, p, tt, prim[i], t[i])
    print("closest-hit mismatches vs oracle: %d / %d  (hit rate %.3f)" % (bad, m, (prim >= 0).mean()))
    oc, light = R.trace_shadow(org, d)
    mats = sc["materials"]
    emit_of = np.array([1 if mm["emit"] else 0 for mm in mats])
    exp = np.where(prim < 0, 0, np.where(emit_of[sc["mat_indices"][np.maximum(prim, 0)]] == 1, 1, 2))
    print("shadow outcome mismatches vs closest-hit rule: %d / %d" % ((exp != oc).sum(), n))
    R.close()
    return bad

def parity(sc, name, spp, w, bvh_kind=0, first=0, count=1):
    g = np.load(os.path.join(G, name + ".npz"))
    sc = dict(sc); sc["width"] = sc["height"] = w
    R = rt.Renderer.from_scene(sc, bvh_kind=bvh_kind)
    t0 = time.time()
    R.render_subframes(first, count, spp)
    dt = time.time() - t0
    acc = R.read_accum()[..., :3]
    st = R.stats()
    if "accum" in g:
        ref = g["accum"]; mine = acc
    else:
        o = g["crop_origin"]; ref = g["crop"]; mine = acc[o[0]:o[0]+128, o[1]:o[1]+128]
    d = np.abs(mine - ref).max(axis=2)
    print("%s: match<1e-4 %.4f  mean mine %s ref %s ratio %s  | %.1f ms  %.2f Msamples/s  %.1f Mrays/s  launches %d rays/sample %.1f" % (
        name, (d < 1e-4).mean(), acc.reshape(-1, 3).mean(0), g["mean_rgb"], acc.reshape(-1, 3).mean(0) / g["mean_rgb"],
        st["last_render_ms"], st["last_samples"] / st["last_render_ms"] / 1e3,
        (st["last_radiance_rays"] + st["last_shadow_rays"]) / st["last_render_ms"] / 1e3, st["last_kernel_launches"],
        (st["last_radiance_rays"] + st["last_shadow_rays"]) / st["last_samples"]))
    R.close()

if __name__ == "__main__":
    sc = load("scenes/cornell_c1.rto")
    which = sys.argv[1:] or ["bvh", "parity", "perf"]
    if "bvh" in which:
        for k in (1, 0):
            print("== bvh kind", k); check_bvh(sc, k)
    if "parity" in which:
        for k in (1, 0):
            print("== parity bvh kind", k)
            parity(sc, "optix_tiny_1", 1, 64, k)
            parity(sc, "optix_tiny_16", 16, 64, k)
            parity(sc, "optix_tiny_4x4", 4, 64, k, 0, 4)
            parity(sc, "optix_c1_1", 1, 512, k)
            parity(sc, "optix_c1", 64, 512, k)
    if "perf" in which:
        for k in (1, 0):
            sc2 = dict(sc); sc2["width"] = sc2["height"] = 2000
            R = rt.Renderer.from_scene(sc2, bvh_kind=k)
            for spp in (4, 16):
                R.reset(); R.render_subframes(0, 1, spp); st = R.stats()
                print("bvh %d 2000x2000 spp %d: %.1f ms, %.2f Msamples/s, %.1f Mrays/s, launches %d iters %d" % (k, spp, st["last_render_ms"],
                      st["last_samples"] / st["last_render_ms"] / 1e3, (st["last_radiance_rays"] + st["last_shadow_rays"]) / st["last_render_ms"] / 1e3,
                      st["last_kernel_launches"], st["iterations"]))
            R.close()
