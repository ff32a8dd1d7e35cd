#!/usr/bin/env python3
"""BASELINE config C5: a 3840x2160, 4096-spp image sample-partitioned over the GPUs of one box.
The 4096 spp are SUBFRAMES subframes of 4096/SUBFRAMES spp (seed = tea<16>(pixel, subframe), the reference's own unit of
independently seeded samples); rank r renders a contiguous block of them over the full image into its own float4 sums and
ONE NCCL reduce puts the total on rank 0.  t_render = first launch -> accumulators final on rank 0, max over ranks,
barrier + synchronize on both sides; parse, upload and BVH build are outside (reported by rank 0).
  python scripts/c5_multi.py [scene.rto [subframes]]                                              (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/c5_multi.py [scene [subframes]]
The image is the same for every N (same subframe set; float additions in a different order)."""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "scenes/cornell_4k.rto"
    subframes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lisa_b200.frontend as fe, lisa_b200.rt as rt
    from lisa_b200 import dist as ldist
    if rank != 0:
        sys.stdout = open(os.devnull, "w")   # the loader's progress lines once, not N times
    sc = fe.parse_scene(scene)
    if len(sys.argv) > 5:
        sc["width"], sc["height"], sc["num_samples"] = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    spp = sc["num_samples"] // subframes
    t0 = time.perf_counter()
    R = rt.Renderer.from_scene(sc, device=local)
    t_create = time.perf_counter() - t0
    R.render_subframes(1000 + rank, 1, 1); R.reset()   # warm-up: module load, allocator, clocks
    if world > 1:
        w = torch.zeros(1 << 20, device="cuda"); dist.all_reduce(w)   # NCCL communicator set-up outside the timed region
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    f, n = ldist.render_partitioned(R, 0, subframes, spp)
    R.sync(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t1], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    st = R.stats()
    mine = torch.tensor([st["last_render_ms"], float(st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"])],
                        device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    if rank == 0:
        img = R.read_accum()[..., :3]
        samples = sc["width"] * sc["height"] * spp * subframes
        print(json.dumps(dict(scene=scene, n_gpus=world, width=sc["width"], height=sc["height"], spp=spp * subframes, subframes=subframes,
                              bounces=sc["num_bounces"], t_render_s=round(float(dt.item()), 3),
                              msamples_per_s=round(samples / float(dt.item()) / 1e6, 1),
                              mrays_traversed_per_s=round(float(mine[1].item()) / float(dt.item()) / 1e6, 1),
                              sum_of_device_render_ms=round(float(mine[0].item()), 1), create_s_rank0=round(t_create, 3),
                              reduce_bytes=int(R.accum_bytes()) if world > 1 else 0,
                              mean_rgb=[float(x) for x in img.reshape(-1, 3).mean(0)])), file=sys.__stdout__, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
