"""The drop-in, compiled: the REFERENCE'S OWN front-end — its scene_parser.cc, parse_obj.cc,
scene_parser.hh / structs.hh / parse_args.hh, compiled where they lie under /root/reference by oracle/Makefile — with
integration/lisa_main_dropin.cc standing in for src/LiSA/src/main.cc:16-24 (OptixWrapper / render / display replaced by
lisa_create / lisa_render_subframes / lisa_write_image) and linked against the product's liblisa_rt.so.  The binary is
built in the build container (oracle/_ref/lisa_dropin; it travels to the GPU box like the other prebuilt files) and
must produce, byte for byte, the PPM the product's own CLI writes."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

DROPIN = os.path.join(ROOT, "oracle", "_ref", "lisa_dropin")
LISA = os.path.join(ROOT, "lisa_b200", "lisa")
HAVE_REF = os.path.isdir("/root/reference/src/LiSA")


def _ensure_built():
    if HAVE_REF:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/lisa_dropin"])
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/lisa_dropin was not prebuilt and the reference sources are not here")


def test_dropin_links_only_the_product_library(built):
    """Builds from the reference's sources + the binding file; needs no OptiX, GL or sutil at link or load time."""
    _ensure_built()
    out = subprocess.run(["ldd", DROPIN], stdout=subprocess.PIPE, text=True).stdout
    assert "liblisa_rt.so" in out
    assert not re.search(r"optix|libGL|glfw|sutil", out, re.I)
    r = subprocess.run([DROPIN], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)  # main.cc:8-13
    assert r.returncode == 1 and r.stderr.splitlines() == ["Missing scene path.", "Usage: %s -s scene_path" % DROPIN]
    r = subprocess.run([DROPIN, "-s", "tests/golden/parser_cases/no_height.rto"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and r.stderr == "Param height not found.\n"   # the reference parser's own error path


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [[], ["-d"]], ids=["render", "display"])
def test_dropin_image_equals_product_cli(built, tmp_path, flags):
    _ensure_built()
    outs = []
    for exe, name in ((DROPIN, "dropin"), (LISA, "product")):
        out = tmp_path / (name + ".ppm")
        txt = open(os.path.join(ROOT, "scenes", "cornell_tiny.rto")).read()
        txt = re.sub(r"output_image\s*=.*", "output_image = %s" % out, txt)
        txt = re.sub(r"num_samples\s*=\s*\d+", "num_samples = 40", txt)   # -d: subframes of 16 spp -> 3 launches (render.cc:121)
        scene = tmp_path / (name + ".rto")
        scene.write_text(txt)
        r = subprocess.run([exe, "-s", str(scene), *flags], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        so = r.stdout.decode()
        assert "Importing assets/objs/cornell_box/bot.obj...\nDone. Imported 2 triangles.\n" in so and "Starting rendering...\n" in so
        outs.append(out.read_bytes())
    assert outs[0].startswith(b"P6\n64 64\n255\n") and len(outs[0]) == 13 + 64 * 64 * 3
    assert outs[0] == outs[1]
