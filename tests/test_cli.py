"""The `lisa` CLI keeps the reference's surface (src/LiSA/src/main.cc:6-28): -s mandatory, usage + exit 1
without it, parser failures print the reference's message and exit with its code."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

LISA = os.path.join(ROOT, "lisa_b200", "lisa")


def _run(*args):
    return subprocess.run([LISA, *args], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)


def test_missing_scene_flag(built):
    r = _run()
    assert r.returncode == 1
    assert r.stderr.decode().splitlines() == ["Missing scene path.", "Usage: %s -s scene_path" % LISA]
    r = _run("-s")  # flag without a value: the reference dereferences a null path; here it is the same usage error
    assert r.returncode == 1


def test_scene_errors_keep_reference_messages(built):
    r = _run("-s", "tests/golden/parser_cases/no_height.rto")
    assert r.returncode == 1 and r.stderr.decode() == "Param height not found.\n"
    r = _run("-s", "tests/golden/parser_cases/missing_obj.rto")
    assert r.returncode == 255 and r.stderr.decode().strip() == "tests/golden/parser_cases/missing.obj not found."
    assert r.stdout.decode() == "Importing tests/golden/parser_cases/missing.obj...\n"
    r = _run("-s", "does_not_exist.rto")
    assert r.returncode == 1 and r.stderr.decode() == "does_not_exist.rto not found\n"


@pytest.mark.gpu
def test_render_end_to_end(built, tmp_path):
    """`lisa -s` renders and writes the PPM named by output_image; `-d` is the progressive mode."""
    import numpy as np
    os.makedirs(os.path.join(ROOT, "out"), exist_ok=True)
    out = os.path.join(ROOT, "out", "cornell_tiny.ppm")
    for flags in ([], ["-d"], ["--split-triangles"]):
        if os.path.exists(out):
            os.remove(out)
        r = _run("-s", "scenes/cornell_tiny.rto", "--stats", *flags)
        assert r.returncode == 0, r.stderr.decode()
        so = r.stdout.decode()
        assert "Importing assets/objs/cornell_box/bot.obj...\nDone. Imported 2 triangles.\n" in so
        assert "Starting rendering...\n" in so and "Rendering finished in " in so and " mn.\n" in so
        assert '"triangles": 1002, "references": 1002,' in so   # --stats; the Cornell walls are not worth splitting
        if flags == ["-d"]:
            assert "nb sample   :       16" in so
        raw = open(out, "rb").read()
        assert raw.startswith(b"P6\n64 64\n255\n") and len(raw) == len(b"P6\n64 64\n255\n") + 64 * 64 * 3
        px = np.frombuffer(raw[13:], dtype=np.uint8).reshape(64, 64, 3)
        assert px[2:12, 20:44].mean() > px[30:, :].mean()  # the light is at the TOP of the file (rows are flipped)


@pytest.mark.gpu
def test_cli_checkpoint_and_resume(built, tmp_path):
    """`-d --snapshot-every K --checkpoint f` rewrites the PPM and saves the accumulators every K
    subframes; `--resume f` continues from the subframe the file holds and ends with the uninterrupted render's image."""
    out_a, out_b, ck = tmp_path / "a.ppm", tmp_path / "b.ppm", tmp_path / "acc.bin"

    def scene(path, out, spp):
        txt = open(os.path.join(ROOT, "scenes/cornell_tiny.rto")).read()
        txt = re.sub(r"output_image\s*=.*", "output_image = %s" % out, txt)
        txt = re.sub(r"num_samples\s*=\s*\d+", "num_samples = %d" % spp, txt)
        path.write_text(txt)
        return str(path)

    def run(*args):
        r = _run(*args)
        assert r.returncode == 0, r.stderr.decode()
        return r.stdout.decode()

    run("-s", scene(tmp_path / "full.rto", out_a, 64), "-d")                                   # 4 subframes straight
    so = run("-s", scene(tmp_path / "half.rto", out_b, 32), "-d", "--snapshot-every", "1", "--checkpoint", str(ck))
    assert so.count("nb sample") == 2 and ck.exists()
    assert ck.stat().st_size == 8 + 12 + 64 * 64 * 16 and ck.read_bytes()[:8] == b"LISAACC1"
    so = run("-s", scene(tmp_path / "rest.rto", out_b, 64), "-d", "--resume", str(ck))         # subframes 2 and 3
    assert "resumed %s at subframe 2 (32 samples)" % ck in so and so.count("nb sample") == 2 and "nb sample   :       64" in so
    assert out_a.read_bytes() == out_b.read_bytes()
    # a checkpoint that already holds every sample: nothing is rendered, the image is written
    so = run("-s", str(tmp_path / "half.rto"), "-d", "--resume", str(ck))
    assert so.count("nb sample") == 0 and "Rendering finished" in so
    # a file that is not a checkpoint is refused
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"P6\n64 64\n255\n" + bytes(100))
    r = _run("-s", str(tmp_path / "rest.rto"), "-d", "--resume", str(bad))
    assert r.returncode != 0 and b"is not an accumulator checkpoint" in r.stderr
