#!/usr/bin/env python3
"""A BASELINE config at FULL spec in one call, like the reference's `-s` (subframe 0, one chain of num_samples per
pixel, PPM written); prints one JSON line.  Default: configs[1], README Cornell box 2000x2000, 2000 spp, 7 bounces.
  full_c2.py [scene.rto [width height spp]]     e.g. scenes/cornell_4k.rto (C5), scenes/c3_knot.rto 3840 2160 4096"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lisa_b200.frontend as fe, lisa_b200.rt as rt
scene = sys.argv[1] if len(sys.argv) > 1 else "scenes/cornell_c2.rto"
sc = fe.parse_scene(scene)
if len(sys.argv) > 4:
    sc["width"], sc["height"], sc["num_samples"] = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
os.makedirs("out", exist_ok=True)
t0 = time.perf_counter()
R = rt.Renderer.from_scene(sc)
R.render()                       # lisa_render_subframes(0, 1, num_samples)
R.write_ppm("out/%s.ppm" % os.path.splitext(os.path.basename(scene))[0])
wall = time.perf_counter() - t0
st = R.stats()
img = R.read_accum()
print(json.dumps(dict(scene=scene, triangles=st["num_triangles"], width=sc["width"], height=sc["height"], spp=sc["num_samples"], bounces=sc["num_bounces"], render_ms=round(st["last_render_ms"], 1),
      wall_s_incl_create_and_ppm=round(wall, 2), msamples_per_s=round(st["last_samples"] / st["last_render_ms"] / 1e3, 1),
      mrays_traversed_per_s=round((st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]) / st["last_render_ms"] / 1e3, 1),
      reference_equivalent_mrays_per_s=round((st["last_radiance_rays"] + st["last_shadow_rays"]) / st["last_render_ms"] / 1e3, 1),
      iterations=st["iterations"], kernel_launches=st["last_kernel_launches"], mean_rgb=[float(x) for x in img[..., :3].reshape(-1, 3).mean(0)])))
