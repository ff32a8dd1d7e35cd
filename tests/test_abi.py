"""The C-ABI libraries load and export every symbol include/*.h declares; argument validation and the
no-GPU behaviour (no compute happens here: the CPU suite only checks that the product FAILS LOUDLY
without a CUDA device instead of falling back)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lisa_[a-z0-9_]+)\s*\(", txt)))


def test_rt_exports_every_declared_symbol(rt):
    names = _declared("lisa_rt.h")
    assert len(names) >= 18
    lib = ctypes.CDLL(rt.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "liblisa_rt.so does not export %s" % n
    assert sorted(rt.EXPORTS) == names


def test_host_exports_every_declared_symbol(frontend):
    names = [n for n in _declared("lisa_host.h")]
    lib = ctypes.CDLL(frontend.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "liblisa_host.so does not export %s" % n
    assert sorted(frontend.EXPORTS) == names


def test_struct_layouts(rt):
    assert ctypes.sizeof(rt.Material) == 40          # structs.hh:16-23
    assert rt.Material.diffuse_color.offset == 12 and rt.Material.emit.offset == 24 and rt.Material.emission_color.offset == 28
    assert ctypes.sizeof(rt.Camera) == 28
    assert rt.library().lisa_version() == 1


def test_argument_validation(rt):
    L = rt.library()
    out = ctypes.c_void_p()
    assert L.lisa_create(None, None, ctypes.byref(out)) == -1
    sd = rt.SceneDesc()
    sd.width, sd.height = 4, 4
    sd.num_vertices = 4  # not a multiple of 3
    assert L.lisa_create(ctypes.byref(sd), None, ctypes.byref(out)) == -1
    assert b"multiple of 3" in L.lisa_last_error()
    sd.num_vertices = 0
    sd.width = 0
    assert L.lisa_create(ctypes.byref(sd), None, ctypes.byref(out)) == -1
    # out-of-range material index is rejected before any device work
    v = np.zeros((3, 3), np.float32)
    m = np.array([3], np.int32)
    mats = rt.pack_materials([dict(alpha=1.0, diffuse=(1, 1, 1), roughness=1.0)])
    sd = rt.SceneDesc()
    sd.vertices, sd.normals, sd.mat_indices = v.ctypes.data, v.ctypes.data, m.ctypes.data
    sd.materials = ctypes.cast(mats, ctypes.c_void_p)
    sd.num_vertices, sd.num_materials, sd.width, sd.height = 3, 1, 2, 2
    assert L.lisa_create(ctypes.byref(sd), None, ctypes.byref(out)) == -1
    assert b"material index 3 out of range" in L.lisa_last_error()
    assert L.lisa_render_subframes(None, 0, 1, 1) == -1
    assert L.lisa_kat_eval(-1, 99, 1, None, None, None, None) == -1


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_fallback(rt, cornell):
    """Without a GPU the product refuses to render: LISA_ERR_CUDA, message says there is no CPU fallback."""
    with pytest.raises(rt.LisaError) as e:
        rt.Renderer.from_scene(cornell)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(rt.LisaError):
        rt.kat_eval(0, in_u=[[0, 0]])


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under lisa_b200/ or include/ may reference it."""
    bad = []
    for base in ("lisa_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cc", ".hh", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|cpu_ref", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_second_bsdf_variant_builds_and_exports(built):
    """The BSDF seam is real: the same kernels compile against bsdf/ggx.cuh (make BSDF=ggx) into a library with the
    same C ABI."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lisa_b200"), "BSDF=ggx", "variant"])
    lib = ctypes.CDLL(os.path.join(ROOT, "lisa_b200", "liblisa_rt_ggx.so"))
    for n in _declared("lisa_rt.h"):
        assert hasattr(lib, n), n


def test_stats_and_options_mirror_the_header(rt, tmp_path):
    """The ctypes mirrors of lisa_stats / lisa_options (lisa_b200/rt.py) against the C header itself: a C program compiled
    from include/lisa_rt.h prints sizeof and the offsets of the fields added last."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lisa_rt.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(lisa_stats), offsetof(lisa_stats, bvh_sah_nodes_per_ray),\n'
                   '  offsetof(lisa_stats, pool_flavour), offsetof(lisa_stats, build_sort_ms), offsetof(lisa_stats, num_references), sizeof(lisa_options)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    size, o_sah, o_flav, o_sort, o_refs, size_opt = map(int, subprocess.check_output([str(exe)]).split())
    S = rt.Stats
    assert ctypes.sizeof(S) == size
    assert S.bvh_sah_nodes_per_ray.offset == o_sah and S.pool_flavour.offset == o_flav
    assert S.build_sort_ms.offset == o_sort and S.num_references.offset == o_refs
    assert ctypes.sizeof(rt.Options) == size_opt
