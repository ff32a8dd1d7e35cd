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


def test_lisa_multi_on_one_gpu_is_the_single_context(rt, cornell):
    """lisa_multi with ONE GPU (runs on any box): same accumulators as a plain context, for subframe blocks and for the
    `-s` split (render_samples: N samples, exactly), and a second call keeps accumulating."""
    sc = resized(cornell, 64)
    M = rt.MultiRenderer.from_scene(sc, num_gpus=1)
    assert M.num_gpus() == 1 and M.backend() == "single"
    R = rt.Renderer.from_scene(sc)
    M.render_subframes(2, 3, 4); R.render_subframes(2, 3, 4)
    np.testing.assert_array_equal(M.read_accum(), R.read_accum())
    M.render_subframes(5, 2, 4); R.render_subframes(5, 2, 4)
    np.testing.assert_array_equal(M.read_accum(), R.read_accum())
    assert M.stats()["samples"] == R.stats()["samples"] == 64 * 64 * 20
    M.reset(); R.reset()
    M.render_samples(0, 7); R.render_subframes(0, 1, 7)
    np.testing.assert_array_equal(M.read_accum(), R.read_accum())
    M.close(); R.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("nccl", ["1", "0"], ids=["nccl", "peer"])
def test_lisa_multi_two_gpus_equal_one(rt, cornell, nccl):
    """NCCL inside the product: one process, one context per GPU, ncclCommInitAll + ONE ncclReduce
    of the float4 accumulators.  The 2-GPU image equals the 1-GPU image over the same subframes (same samples, another
    order of the float additions); samples are counted exactly when G does not divide N; LISA_NCCL=0 forces the
    peer-kernel reduce and must give the same picture.  Run in a subprocess: the reduce backend is chosen once per process."""
    code = ("import os, sys, numpy as np\n"
            "import lisa_b200.frontend as fe, lisa_b200.rt as rt\n"
            "sc = fe.parse_scene('scenes/cornell_c1.rto'); sc['width'] = sc['height'] = 96\n"
            "M = rt.MultiRenderer.from_scene(sc, num_gpus=2); R = rt.Renderer.from_scene(sc, device=0)\n"
            "print('backend', M.backend(), M.num_gpus())\n"
            "M.render_subframes(3, 5, 4); R.render_subframes(3, 5, 4)\n"
            "np.testing.assert_allclose(M.read_accum(), R.read_accum(), rtol=2e-6, atol=1e-7)\n"
            "M.render_subframes(8, 2, 4); R.render_subframes(8, 2, 4)       # accumulates across calls; the non-root buffers were cleared\n"
            "np.testing.assert_allclose(M.read_accum(), R.read_accum(), rtol=2e-6, atol=1e-7)\n"
            "assert M.stats()['samples'] == R.stats()['samples'] == 96 * 96 * 28 and M.stats()['subframes_accumulated'] == 7\n"
            "M.reset(); R.reset()\n"
            "M.render_samples(0, 7)                                          # 4 + 3 samples: exactly 7, not 2 x ceil(7 / 2)\n"
            "R.render_subframes(0, 1, 4); R.render_subframes(1, 1, 3)\n"
            "np.testing.assert_allclose(M.read_accum(), R.read_accum(), rtol=2e-6, atol=1e-7)\n"
            "assert M.stats()['samples'] == 96 * 96 * 7\n"
            "print('times', M.last_times()); print('OK')\n")
    env = dict(os.environ, LISA_NCCL=nccl)
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:]
    assert ("backend nccl 2" if nccl == "1" else "backend peer 2") in r.stdout, r.stdout[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_chunked_upload_on_every_device(rt, monkeypatch):
    """The pinned upload ring serves every device of the process (its events used to belong to the first device
    that used it, so a large upload to any other GPU failed)."""
    monkeypatch.setenv("LISA_UPLOAD_CHUNKED_MIN", "0")
    rng = np.random.default_rng(5)
    T = 400_000
    c = rng.random((T, 1, 3), dtype=np.float32)
    v = (c + (rng.random((T, 3, 3), dtype=np.float32) - 0.5) * np.float32(0.01)).reshape(-1, 3)
    n = np.repeat(np.float32([[0, 1, 0]]), 3 * T, axis=0)
    m = np.zeros(T, np.int32)
    mats = [dict(emit=False, alpha=1.0, diffuse=(0.8, 0.8, 0.8), roughness=1.0)]
    o = rng.uniform(0, 1, size=(20000, 3)).astype(np.float32)
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    res = []
    for dev in (1, 0, 1):
        R = rt.Renderer(v, n, m, mats, 8, 8, (0, 0, 5), (0, 0, 0), 45.0, 1, 3, device=dev)
        res.append(R.trace_closest(o, d))
        R.close()
    for r in res[1:]:
        np.testing.assert_array_equal(res[0][0], r[0])
        np.testing.assert_array_equal(res[0][1], r[1])


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_cli_gpus_flag_and_peer_reduce(rt, frontend, built):
    """`lisa -s scene --gpus 2` (one process, one context per GPU, ONE ncclReduce through lisa_multi)
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
