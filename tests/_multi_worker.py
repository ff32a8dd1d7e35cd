"""Worker for test_gpu_multi.py: one rank per GPU under torchrun, sample-space partition + one NCCL reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)


def main():
    out, first, count, spp, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lisa_b200.frontend as fe
    import lisa_b200.rt as rt
    from lisa_b200 import dist as ldist
    sc = fe.parse_scene("scenes/cornell_c1.rto")
    sc["width"] = sc["height"] = size
    R = rt.Renderer.from_scene(sc, device=local)
    f, n = ldist.render_partitioned(R, first, count, spp)
    if rank == 0:
        np.save(out, R.read_accum())   # accumulators now hold the sum over ALL ranks
    stats = R.stats()
    print("rank %d rendered subframes [%d, %d) on cuda:%d, %d samples" % (rank, f, f + n, R.device(), stats["samples"]), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
