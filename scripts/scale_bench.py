#!/usr/bin/env python3
"""BASELINE configs C3 (870k-triangle dielectric mesh in the Cornell box) and C4 (triangle soups of 1M/10M/100M):
BVH build time, node/triangle counts, traversal throughput, and a closest-hit check against the oracle.
  python scripts/scale_bench.py knot [spp]      |  python scripts/scale_bench.py soup 1000000 [spp]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)
import lisa_b200.frontend as fe, lisa_b200.rt as rt

def report(name, R, w, h, spp, t_create, extra=None):
    R.render_subframes(100, 1, 1); R.reset()      # warm-up
    R.render_subframes(0, 1, spp)
    st = R.stats()
    rays = st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]
    out = dict(scene=name, triangles=st["num_triangles"], emitter_triangles=st["num_emitter_triangles"], bvh_nodes=st["bvh_nodes"],
               bvh_MB=round(st["bvh_bytes"] / 1e6, 2), triangle_MB=round(st["triangle_bytes"] / 1e6, 2), upload_ms=round(st["upload_ms"], 2),
               bvh_build_ms=round(st["bvh_build_ms"], 2), create_wall_ms=round(t_create * 1e3, 1), width=w, height=h, spp=spp,
               render_ms=round(st["last_render_ms"], 2), msamples_per_s=round(st["last_samples"] / st["last_render_ms"] / 1e3, 2),
               mrays_traversed_per_s=round(rays / st["last_render_ms"] / 1e3, 1),
               reference_rays_per_sample=round((st["last_radiance_rays"] + st["last_shadow_rays"]) / st["last_samples"], 2),
               traversed_rays_per_sample=round(rays / st["last_samples"], 2),
               nodes_per_ray=round(st["last_nodes_visited"] / max(rays, 1), 2), tris_per_ray=round(st["last_triangles_tested"] / max(rays, 1), 2))
    if extra: out.update(extra)
    print(json.dumps(out), flush=True)
    return out

def check_vs_oracle(R, verts, normals, midx, mats_packed, n=4000):
    from oracle import binding
    S = binding.Scene(verts, normals, midx, mats_packed)
    rng = np.random.default_rng(0)
    lo, hi = verts.min(0), verts.max(0)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    prim, t = R.trace_closest(o, d)
    bad = 0
    for i in range(n):
        p, tt = S.closest_hit(o[i], d[i])
        if p != prim[i] and not (p >= 0 and prim[i] >= 0 and abs(tt - t[i]) <= 2e-5 * max(1, abs(tt))): bad += 1
    return bad, n

def main():
    kind = sys.argv[1]
    if kind == "knot":
        spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
        sys.path.insert(0, os.path.join(ROOT, "assets"))
        import gen_knot
        sc = fe.parse_scene("scenes/cornell_c1.rto")
        kv, kn = gen_knot.soup_arrays()
        # Cornell walls + light (drop the two blocks and the sphere), plus the knot in glass (material 1: n = 1.5)
        keep = np.isin(np.arange(len(sc["mat_indices"])), np.r_[0:16, len(sc["mat_indices"]) - 2:len(sc["mat_indices"])])
        v = np.concatenate([sc["vertices"].reshape(-1, 3, 3)[keep].reshape(-1, 3), kv])
        n = np.concatenate([sc["normals"].reshape(-1, 3, 3)[keep].reshape(-1, 3), kn])
        m = np.concatenate([sc["mat_indices"][keep], np.full(len(kv) // 3, 1, np.int32)])
        w, h = 1920, 1080
        t0 = time.perf_counter()
        R = rt.Renderer(v, n, m, sc["materials_packed"], w, h, sc["camera"]["eye"], sc["camera"]["look_at"], sc["camera"]["fov"], spp, 12)
        tc = time.perf_counter() - t0
        bad, nn = check_vs_oracle(R, v, n, m, sc["materials_packed"], 3000)
        report("C3 knot 871k glass in Cornell, 12 bounces", R, w, h, spp, tc, {"closest_hit_mismatches_vs_oracle": "%d/%d" % (bad, nn)})
        os.makedirs("out", exist_ok=True); R.write_ppm("out/c3_knot.ppm")
    else:
        T = int(float(sys.argv[2])); spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
        rng = np.random.default_rng(0x5EED)
        edge = 0.5 * T ** (-1.0 / 3.0)
        c = rng.random((T, 1, 3), dtype=np.float32)
        v = (c + (rng.random((T, 3, 3), dtype=np.float32) - 0.5) * np.float32(2 * edge)).reshape(-1, 3)
        e1 = v[1::3] - v[0::3]; e2 = v[2::3] - v[0::3]
        fn = np.cross(e1, e2); fn /= (np.linalg.norm(fn, axis=1, keepdims=True) + 1e-30)
        n = np.repeat(fn.astype(np.float32), 3, axis=0)
        # one emissive quad above the cube
        q = np.float32([[-0.5, 1.6, -0.5], [1.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, 1.5]])
        v = np.concatenate([v, q]); n = np.concatenate([n, np.tile(np.float32([[0, -1, 0]]), (6, 1))])
        m = np.concatenate([np.zeros(T, np.int32), np.ones(2, np.int32)])
        mats = [dict(emit=False, alpha=1.0, diffuse=(0.7, 0.7, 0.7), roughness=1.0), dict(emit=True, alpha=1.0, emission=(1, 1, 1))]
        w = h = 1024
        t0 = time.perf_counter()
        R = rt.Renderer(v, n, m, mats, w, h, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, spp, 7)
        tc = time.perf_counter() - t0
        extra = {}
        if T <= 2_000_000:
            from oracle.scene_py import pack_material
            mp = b"".join(pack_material(roughness=x.get("roughness", 0), alpha=x["alpha"], diffuse=x.get("diffuse", (0, 0, 0)), emit=x["emit"], emission=x.get("emission", (0, 0, 0))) for x in mats)
            bad, nn = check_vs_oracle(R, v, n, m, mp, 2000)
            extra["closest_hit_mismatches_vs_oracle"] = "%d/%d" % (bad, nn)
        report("C4 soup %d" % T, R, w, h, spp, tc, extra)

if __name__ == "__main__":
    main()
