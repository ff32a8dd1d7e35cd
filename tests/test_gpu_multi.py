"""Sample-space partition over 2 GPUs: N ranks, each a full scene + BVH and a disjoint block of
subframes, ONE NCCL reduce of the float4 sums — must equal the 1-GPU render of the same subframe set (up to the
order of the float additions: rtol 2e-6).  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, resized

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_ranks_equal_one(rt, cornell, tmp_path):
    first, count, spp, size = 3, 5, 4, 96
    out = str(tmp_path / "multi.npy")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "_multi_worker.py"), out, str(first), str(count), str(spp), str(size)]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    multi = np.load(out)
    R = rt.Renderer.from_scene(resized(cornell, size))
    R.render_subframes(first, count, spp)
    single = R.read_accum()
    np.testing.assert_allclose(multi, single, rtol=2e-6, atol=1e-7)
    assert "rank 1 rendered subframes [6, 8)" in r.stdout and "rank 0 rendered subframes [3, 6)" in r.stdout


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_cli_gpus_flag_and_peer_reduce(rt, frontend, built):
    """`lisa -s scene --gpus 2` (one process, one context per GPU, reduce = ONE kernel reading peer memory over NVLink)
    writes the same PPM as a single GPU rendering the same two subframes."""
    lisa = os.path.join(ROOT, "lisa_b200", "lisa")
    os.makedirs(os.path.join(ROOT, "out"), exist_ok=True)
    out = os.path.join(ROOT, "out", "cornell_tiny.ppm")
    if os.path.exists(out):
        os.remove(out)
    r = subprocess.run([lisa, "-s", "scenes/cornell_tiny.rto", "--gpus", "2"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    raw = open(out, "rb").read()
    sc = frontend.parse_scene("scenes/cornell_tiny.rto")
    R = rt.Renderer.from_scene(sc)
    R.render_subframes(0, 2, 8)          # num_samples = 16 -> 2 subframes of 8 spp
    px = R.read_rgba8()
    img = np.frombuffer(raw[len(b"P6\n64 64\n255\n"):], np.uint8).reshape(64, 64, 3)
    np.testing.assert_array_equal(img, px[::-1, :, :3])
    # the same through the binding: two contexts on two devices, accum_add_peer
    A = rt.Renderer.from_scene(sc, device=0)
    B = rt.Renderer.from_scene(sc, device=1)
    A.render_subframes(0, 1, 8)
    B.render_subframes(1, 1, 8)
    A.accum_add_peer(B)
    np.testing.assert_array_equal(A.read_accum(), R.read_accum())
