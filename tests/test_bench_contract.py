"""bench.py's output contract on a box without a GPU: `--impl reference` falls back from the OptiX harness to the CPU
restatement (labelled "port"), and stdout carries exactly ONE JSON line whatever the libraries print (the OBJ loader's
progress lines go to stderr); the product arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


@pytest.mark.skipif(_has_cuda(), reason="exercises the no-GPU paths")
def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, LISA_BENCH_CPU_TARGET_S="2")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = r.stdout.decode().splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reference arm neither imports the product package nor maps its libraries (the config comes from the scene text)
    assert "product libraries mapped in this process: []; product modules imported: []" in r.stderr.decode()
    # both arms name the workload with the same dictionary
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_dict(2000, 2000, 7, 50) and bench.scene_header() == (2000, 2000, 7)


@pytest.mark.skipif(_has_cuda(), reason="exercises the no-GPU paths")
def test_product_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0"], cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    assert r.returncode != 0 and r.stdout.decode().strip() == ""
    assert "no CUDA device" in r.stderr.decode()
