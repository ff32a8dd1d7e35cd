#!/usr/bin/env python3
"""Developer A/B on the GPU box: scripts/ab_variants.py [--scenes c2,c3,soup1m] [--rounds 2] NAME[=ENV=VAL,...] ...
Each NAME is a library built by scripts/mkvariant.sh (lisa_b200/variants/liblisa_rt_NAME.so; `product` = the in-tree
liblisa_rt.so).  Every (variant, scene) runs in its own process, variants interleaved per round; prints the best launch time,
Msamples/s and a hash of the accumulators (variants of one scene must agree: the kernels are bit-reproducible)."""
import hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = {
    "c2": ("scenes/cornell_c2.rto", 2000, 2000, 50, 4),
    "c1": ("scenes/cornell_c1.rto", 512, 512, 64, 4),
    "c3": ("scenes/c3_knot.rto", 1920, 1080, 16, 4),
    "c256": ("scenes/cornell_c1.rto", 256, 256, 64, 4),
    "c720p": ("scenes/cornell_c1.rto", 1280, 720, 16, 4),
    "c1024": ("scenes/cornell_c1.rto", 1024, 1024, 16, 4),
    "c3s": ("scenes/c3_knot.rto", 480, 270, 64, 4),
}

def child(scene, reps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import lisa_b200.frontend as fe
    import lisa_b200.rt as rt
    if scene.startswith("soup"):  # the C4 soup of scripts/scale_bench.py
        T = int(float(scene[4:].replace("m", "e6").replace("k", "e3")))
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
        w = h = 1024; spp = 4
        R = rt.Renderer(v, n, m, mats, w, h, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, spp, 7)
    else:
        path, w, h, spp, _ = SCENES[scene]
        sc = fe.parse_scene(os.path.join(ROOT, path))
        sc["width"], sc["height"] = w, h
        R = rt.Renderer.from_scene(sc)
    R.render_subframes(0, 1, 2)
    ms = []
    for i in range(reps):
        R.render_subframes(1 + i, 1, spp)
        ms.append(R.stats()["last_render_ms"])
    acc = R.read_accum()
    st = R.stats()
    R.close()
    print(json.dumps({"ms": min(ms), "all_ms": ms, "msamples": w * h * spp / min(ms) / 1e3, "hash": hashlib.md5(np.ascontiguousarray(acc).tobytes()).hexdigest()[:10],
                      "nodes": st.get("last_nodes_visited"), "tris": st.get("last_triangles_tested"),
                      "sah": st.get("bvh_sah_nodes_per_ray"), "flavour": st.get("pool_flavour"), "bvh_nodes": st.get("bvh_nodes"), "build_ms": st.get("bvh_build_ms"),
                      "rays": st.get("last_radiance_rays", 0) + st.get("last_shadow_rays", 0) - st.get("last_shadow_culled", 0)}))

def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        return child(args[1], int(args[2]))
    scenes, rounds, reps, variants = ["c2"], 2, 4, []
    while args:
        a = args.pop(0)
        if a == "--scenes": scenes = args.pop(0).split(",")
        elif a == "--rounds": rounds = int(args.pop(0))
        elif a == "--reps": reps = int(args.pop(0))
        else: variants.append(a)
    best = {}
    for r in range(rounds):
        for v in variants:
            name, _, envs = v.partition("=")
            env = dict(os.environ)
            lib = name.split("+")[0]
            if lib != "product": env["LISA_RT_LIB"] = os.path.join(ROOT, "lisa_b200", "variants", "liblisa_rt_%s.so" % lib)
            for kv in (envs.split(",") if envs else []):
                k, _, val = kv.partition(":")
                env[k] = val
            for s in scenes:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", s, str(reps)], env=env, capture_output=True, text=True)
                try: o = json.loads(p.stdout.strip().splitlines()[-1])
                except Exception:
                    print("%-22s %-7s FAILED: %s" % (v, s, (p.stderr or p.stdout)[-400:]), flush=True); continue
                print("round %d %-22s %-7s %8.2f ms %8.1f Msamples/s hash %s  sah %s flavour %s nodes %s build %.2f ms  %.2f nodes %.2f tris per ray" % (r, v, s, o["ms"], o["msamples"], o["hash"], o.get("sah"), o.get("flavour"), o.get("bvh_nodes"), o.get("build_ms") or 0, (o.get("nodes") or 0) / max(o.get("rays") or 1, 1), (o.get("tris") or 0) / max(o.get("rays") or 1, 1)), flush=True)
                k = (v, s)
                if k not in best or o["ms"] < best[k]["ms"]: best[k] = o
    print("---- best of %d rounds" % rounds)
    for s in scenes:
        base = best.get((variants[0], s))
        for v in variants:
            o = best.get((v, s))
            if o: print("%-22s %-7s %8.2f ms %8.1f Msamples/s  x%.4f  hash %s" % (v, s, o["ms"], o["msamples"], base["ms"] / o["ms"] if base else 0, o["hash"]))

if __name__ == "__main__":
    main()
