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
