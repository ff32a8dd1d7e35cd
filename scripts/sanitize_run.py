"""Small end-to-end render for compute-sanitizer (memcheck / racecheck / initcheck): both BVH kinds, culling on/off,
tiled state, readback."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lisa_b200.frontend as fe, lisa_b200.rt as rt
sc = fe.parse_scene("scenes/cornell_c1.rto")
sc["width"], sc["height"] = 48, 40
for bvh in (0, 1):
    for flags in (0, rt.FLAG_NO_CULL):
        R = rt.Renderer.from_scene(sc, bvh_kind=bvh, flags=flags, max_chains=1500)
        R.render_subframes(0, 2, 3)
        img = R.read_accum(); px = R.read_rgba8()
        print(bvh, flags, float(img[..., :3].mean()), R.stats()["last_kernel_launches"])
        R.close()
# round 2: every persistent schedule, the 20k-triangle soup (several PLOC rounds with the look-back scan, two sort tiles),
# a non-flat emitter set (cone filter) and the serialised BVH
import numpy as np
rng = np.random.default_rng(3)
T = 20000
c = rng.uniform(-1, 1, size=(T, 1, 3))
v = (c + rng.normal(scale=0.05, size=(T, 3, 3))).astype(np.float32).reshape(-1, 3)
n = rng.normal(size=(3 * T, 3)).astype(np.float32); n /= np.linalg.norm(n, axis=1, keepdims=True)
m = (rng.random(T) < 0.02).astype(np.int32)
mats = [dict(emit=False, alpha=1.0, diffuse=(0.8, 0.8, 0.8), roughness=1.0), dict(emit=True, alpha=1.0, emission=(1, 1, 1))]
for pipe in ("pool", "path", "wavefront"):
    os.environ["LISA_PIPELINE"] = pipe
    R = rt.Renderer(v, n, m, mats, 40, 32, (0, 0, 4), (0, 0, 0), 45.0, 1, 3)
    R.render_subframes(0, 1, 2)
    print("soup", pipe, float(R.read_accum()[..., :3].mean()), R.stats()["bvh_nodes"])
    if pipe == "pool":
        R.save_bvh("/tmp/sanitize.lisabvh")
    R.close()
    R = rt.Renderer.from_scene(sc)                      # flat quad light: the flat-bounds filter
    R.render_subframes(0, 1, 2)
    print("cornell", pipe, float(R.read_accum()[..., :3].mean()))
    R.close()
# round 2, second session: the deep flavour of k_pool (48 chains per warp, 12 stack entries) on the soup and on the Cornell box,
# and the collapse with and without the surface-area rule for leaves
os.environ["LISA_PIPELINE"] = "pool"
for fl, sah in (("deep", "0.5"), ("shallow", "-1")):
    os.environ["LISA_POOL_FLAVOUR"] = fl; os.environ["LISA_LEAF_SAH"] = sah
    R = rt.Renderer(v, n, m, mats, 40, 32, (0, 0, 4), (0, 0, 0), 45.0, 1, 3)
    R.render_subframes(0, 1, 2)
    print("soup pool", fl, "leaf_sah", sah, float(R.read_accum()[..., :3].mean()), R.stats()["bvh_nodes"], R.stats()["pool_flavour"])
    R.close()
    R = rt.Renderer.from_scene(sc)
    R.render_subframes(0, 1, 2)
    print("cornell pool", fl, float(R.read_accum()[..., :3].mean()))
    R.close()
os.environ.pop("LISA_POOL_FLAVOUR"); os.environ.pop("LISA_LEAF_SAH")
os.environ.pop("LISA_PIPELINE")
R = rt.Renderer(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.int32), mats, 40, 32, (0, 0, 4), (0, 0, 0), 45.0, 1, 3,
                _bvh_file="/tmp/sanitize.lisabvh")
R.render_subframes(0, 1, 1)
print("from file", float(R.read_accum()[..., :3].mean()))
R.close()
