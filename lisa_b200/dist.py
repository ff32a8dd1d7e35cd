"""Sample-space partitioning over the GPUs of one box — torch.distributed plumbing only.

Every rank holds the full scene and BVH (built locally from the same upload: the builder is
deterministic), renders a disjoint set of subframes over the FULL image into its own float4 sums, and
one reduce over NCCL/NVLink combines them on the root.  Subframes are the reference's own unit of
independently seeded samples (seed = tea<16>(pixel, subframe), shader.cu:140-141), so an N-rank image
equals the 1-rank image over the same subframe set up to the order of the float additions.
"""
import torch
import torch.distributed as dist


def subframes_for_rank(first, count, rank, world):
    """Contiguous block partition of subframes [first, first+count) — returns (first_r, count_r).

    Blocks (not round-robin) so that each rank issues ONE lisa_render_subframes call whose chains all
    run concurrently; the block sizes differ by at most one.
    """
    base, rem = divmod(count, world)
    n = base + (1 if rank < rem else 0)
    f = first + rank * base + min(rank, rem)
    return f, n


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can alias the library's accumulator buffer."""

    def __init__(self, ptr, nfloat):
        self.__cuda_array_interface__ = {"shape": (nfloat,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def accum_tensor(renderer):
    """Zero-copy torch view (float32, W*H*4) of the context's accumulator sums on its device."""
    n = renderer.accum_bytes() // 4
    dev = torch.device("cuda", renderer.device())
    return torch.as_tensor(_CudaArray(renderer.accum_device_ptr(), n), device=dev)


def reduce_accum(t, dst=0, group=None):
    """Sum the per-rank accumulators onto `dst` (one collective: ncclReduce over NVLink, gloo on CPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return t


def render_partitioned(renderer, first, count, spp, group=None):
    """Render this rank's share of subframes and reduce the sums onto rank 0.  Returns (first_r, count_r)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    f, n = subframes_for_rank(first, count, rank, world)
    if n:
        renderer.render_subframes(f, n, spp)
    if world > 1:
        t = accum_tensor(renderer)
        renderer.sync()
        torch.cuda.current_stream(t.device).synchronize()
        reduce_accum(t, 0, group)
        torch.cuda.current_stream(t.device).synchronize()
    return f, n
