#!/usr/bin/env python3
"""BASELINE configs[1] at FULL spec in one call, like the reference's `-s`: README Cornell box 2000x2000, 2000 spp,
7 bounces, subframe 0 (one chain of 2000 samples per pixel), PPM written.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lisa_b200.frontend as fe, lisa_b200.rt as rt
sc = fe.parse_scene("scenes/cornell_c2.rto")
os.makedirs("out", exist_ok=True)
t0 = time.perf_counter()
R = rt.Renderer.from_scene(sc)
R.render()                       # lisa_render_subframes(0, 1, num_samples)
R.write_ppm("out/cornell_c2.ppm")
wall = time.perf_counter() - t0
st = R.stats()
img = R.read_accum()
print(json.dumps(dict(width=sc["width"], height=sc["height"], spp=sc["num_samples"], bounces=sc["num_bounces"], render_ms=round(st["last_render_ms"], 1),
      wall_s_incl_create_and_ppm=round(wall, 2), msamples_per_s=round(st["last_samples"] / st["last_render_ms"] / 1e3, 1),
      mrays_traversed_per_s=round((st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]) / st["last_render_ms"] / 1e3, 1),
      reference_equivalent_mrays_per_s=round((st["last_radiance_rays"] + st["last_shadow_rays"]) / st["last_render_ms"] / 1e3, 1),
      iterations=st["iterations"], kernel_launches=st["last_kernel_launches"], mean_rgb=[float(x) for x in img[..., :3].reshape(-1, 3).mean(0)])))
