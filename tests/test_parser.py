"""Scene front-end: the product's C++ SceneParser/parse_obj (lisa_b200/host, through include/lisa_host.h)
and the oracle's Python restatement, both against dumps of the reference's own parser
(tests/golden/ref_parse_*.json, produced by oracle/ref_parse_main.cc which compiles the reference's
scene_parser.cc + parse_obj.cc in place).  Covers the grammar quirks listed in SURVEY.md §8a-P."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

CASES = sorted(glob.glob(os.path.join(GOLDEN, "ref_parse_*.json")))


def _scene_path(name):
    for d in ("scenes", os.path.join("tests", "golden", "parser_cases")):
        p = os.path.join(d, name + ".rto")
        if os.path.exists(os.path.join(ROOT, p)):
            return p
    raise FileNotFoundError(name)


def _check_against(ref, got):
    for k in ("width", "height", "num_samples", "num_bounces", "output_image"):
        assert got[k] == ref[k], k
    for k in ("eye", "look_at"):
        np.testing.assert_array_equal(np.float32(got["camera"][k]), np.float32(ref["camera"][k]))
    assert np.float32(got["camera"]["fov"]) == np.float32(ref["camera"]["fov"])
    assert got["vertices"].shape[0] == ref["num_vertices"]
    assert len(got["materials"]) == ref["num_materials"]
    for gm, rm in zip(got["materials"], ref["materials"]):
        assert bool(gm["emit"]) == rm["emit"]
        assert np.float32(gm["alpha"]) == np.float32(rm["alpha"])
        for k in ("emission", "diffuse"):
            if k in rm:
                np.testing.assert_allclose(np.float32(gm[k]), np.float32(rm[k]), rtol=1e-7)
        for k in ("n", "roughness"):
            if k in rm:
                assert np.float32(gm[k]) == np.float32(rm[k])
    # geometry: run-length material indices, checksums and the first triangles verbatim
    mi, rle = got["mat_indices"], []
    i = 0
    while i < len(mi):
        j = i
        while j < len(mi) and mi[j] == mi[i]:
            j += 1
        rle.append([int(mi[i]), j - i])
        i = j
    assert rle == ref["mat_indices_rle"]
    v, n = got["vertices"].reshape(-1).astype(np.float64), got["normals"].reshape(-1).astype(np.float64)
    assert abs(v.sum() - ref["vertex_sum"]) <= 1e-9 * max(1, abs(ref["vertex_sum"])) + 1e-9
    assert abs(n.sum() - ref["normal_sum"]) <= 1e-9 * max(1, abs(ref["normal_sum"])) + 1e-9
    wv = (v * ((np.arange(v.size) % 251) + 1)).sum()
    assert abs(wv - ref["vertex_wsum"]) <= 1e-9 * max(1, abs(ref["vertex_wsum"])) + 1e-9
    k = len(ref["first_vertices"])
    np.testing.assert_array_equal(np.float32(v[:k]), np.float32(ref["first_vertices"]))
    np.testing.assert_array_equal(np.float32(n[:k]), np.float32(ref["first_normals"]))


@pytest.mark.parametrize("case", CASES, ids=[os.path.basename(c)[10:-5] for c in CASES])
def test_product_parser_matches_reference(case, frontend, capfd):
    ref = json.load(open(case))
    path = _scene_path(os.path.basename(case)[10:-5])
    if "error" in ref:
        with pytest.raises(frontend.SceneError) as e:
            frontend.parse_scene(path)
        assert str(e.value).strip() == ref["error"].strip()
        assert (e.value.code & 0xff) == ref["exit"]
    else:
        _check_against(ref, frontend.parse_scene(path))


@pytest.mark.parametrize("case", CASES, ids=[os.path.basename(c)[10:-5] for c in CASES])
def test_oracle_parser_matches_reference(case):
    from oracle import scene_py
    ref = json.load(open(case))
    path = _scene_path(os.path.basename(case)[10:-5])
    if "error" in ref:
        with pytest.raises(scene_py.SceneError) as e:
            scene_py.parse_scene(path)
        assert str(e.value).strip() == ref["error"].strip()
        assert (e.value.code & 0xff) == ref["exit"]
    else:
        _check_against(ref, scene_py.parse_scene(path))


def test_readme_scene_verbatim(frontend):
    """The README scene text (README.md:47-131 of the reference), unmodified."""
    sc = frontend.parse_scene(os.path.join("tests", "golden", "readme_scene.rto"), load_meshes=False)
    assert (sc["width"], sc["height"], sc["num_samples"], sc["num_bounces"]) == (2000, 2000, 2000, 7)
    assert sc["output_image"] == "../../images/cornel_box.ppm"
    assert len(sc["materials"]) == 5 and len(sc["mesh_files"]) == 9
    assert sc["mesh_files"][6] == ("assets/objs/cornell_box/small_box.obj", 1)
    assert sc["materials"][1]["n"] == 1.5 and sc["materials"][1]["alpha"] == 0.0
    assert sc["materials"][4]["emit"] and tuple(sc["materials"][4]["emission"]) == (1.0, 1.0, 1.0)


def test_missing_scene_file(frontend):
    with pytest.raises(frontend.SceneError) as e:
        frontend.parse_scene("no/such/file.rto")
    assert "no/such/file.rto not found" in str(e.value) and e.value.code == 1


def test_remove_comments_large_input(frontend, tmp_path):
    """The reference's comment regex recurses per character; the scanner must take megabytes."""
    p = tmp_path / "big.rto"
    body = "/*" + "x\n" * 2_000_000 + "*/\n" + open("tests/golden/parser_cases/no_mesh_no_eol.rto").read()
    p.write_text(body)
    sc = frontend.parse_scene(str(p))
    assert sc["num_samples"] == 7 and sc["camera"]["fov"] == 33.25


def test_quad_truncated_and_indices(frontend):
    sc = frontend.parse_scene("tests/golden/parser_cases/color255.rto")
    v = sc["vertices"].reshape(-1, 3, 3)
    assert v.shape[0] == 2
    np.testing.assert_array_equal(v[1], [[0, 0, 0], [1, 0, 0], [1, 1, 0.5]])  # 4th vertex of the quad dropped
    np.testing.assert_array_equal(sc["normals"].reshape(-1, 3, 3)[0], [[0, 0, 1], [0, 0, 1], [0, 1, 0]])


def test_parallel_obj_loader_matches_oracle(frontend, tmp_path, monkeypatch):
    """The chunked multi-threaded OBJ loader (host/parse_obj.cc) against the oracle's line-by-line restatement of
    parse_obj.cc, with the chunk boundaries forced inside a small file."""
    import sys
    from oracle import scene_py
    sys.path.insert(0, os.path.join(ROOT, "assets"))
    import gen_knot
    obj = tmp_path / "k.obj"
    sys.argv = ["gen_knot", str(obj), "60", "16"]
    gen_knot.main()
    with open(obj, "a") as f:   # lines the loader must ignore or truncate
        f.write("vt 0.5 0.5\n# comment\n v 9 9 9\ng grp\nf 1//1 2//2 3//3 4//4\n")
    rto = tmp_path / "k.rto"
    rto.write_text("material g { color = (1, 1, 1, 0)\n n = 1.5 }\nmesh { obj_file = %s\n material = g }\n"
                   "camera { position = (0,0,1) look_at = (0,0,0) fov = 40 }\nnum_samples = 1\nnum_bounces = 1\n"
                   "width = 4\nheight = 4\noutput_image = x.ppm\n" % obj)
    ref = scene_py.parse_scene(str(rto))
    for th in ("1", "3", "7"):
        monkeypatch.setenv("LISA_OBJ_THREADS", th)
        got = frontend.parse_scene(str(rto))
        np.testing.assert_array_equal(got["vertices"], ref["vertices"])
        np.testing.assert_array_equal(got["normals"], ref["normals"])
        np.testing.assert_array_equal(got["mat_indices"], ref["mat_indices"])
    assert got["vertices"].shape[0] == 3 * (2 * 60 * 16 + 1)


def test_obj_soup_cache(frontend, tmp_path, capfd):
    """Opt-in binary soup cache of the OBJ loader: written after a text parse, read back when the
    OBJ is unchanged (same arrays, same two progress lines), ignored when stale or damaged, never written when off."""
    obj = tmp_path / "t.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvn 1 0 0\nf 1//1 2//1 3//1\nf 1//2 3//2 4//2\n")
    rto = tmp_path / "t.rto"
    rto.write_text("material g { color = (1, 1, 1, 1)\n roughness = 1 }\nmesh { obj_file = %s\n material = g }\n"
                   "mesh { obj_file = %s\n material = g }\n"
                   "camera { position = (0,0,1) look_at = (0,0,0) fov = 40 }\nnum_samples = 1\nnum_bounces = 1\n"
                   "width = 4\nheight = 4\noutput_image = x.ppm\n" % (obj, obj))
    cache = tmp_path / "t.obj.lisasoup"
    plain = frontend.parse_scene(str(rto))
    assert not cache.exists()                      # off by default: nothing is written next to the inputs
    frontend.set_obj_cache(True)
    try:
        capfd.readouterr()
        first = frontend.parse_scene(str(rto))     # mesh 1 parses text and writes the cache, mesh 2 already reads it
        out = capfd.readouterr().out
        assert out == ("Importing %s...\nDone. Imported 2 triangles.\n" % obj) * 2
        assert cache.exists() and cache.stat().st_size == 48 + 2 * 72
        again = frontend.parse_scene(str(rto))
        for k in ("vertices", "normals", "mat_indices"):
            np.testing.assert_array_equal(first[k], plain[k])
            np.testing.assert_array_equal(again[k], plain[k])
        # the cache is really what is read: plant different numbers in it, keeping the header
        raw = bytearray(cache.read_bytes())
        raw[48:52] = np.float32(7.0).tobytes()
        cache.write_bytes(bytes(raw))
        assert frontend.parse_scene(str(rto))["vertices"][0, 0] == 7.0
        # a truncated file is ignored (and rewritten from the text)
        cache.write_bytes(bytes(raw[:100]))
        np.testing.assert_array_equal(frontend.parse_scene(str(rto))["vertices"], plain["vertices"])
        assert cache.stat().st_size == 48 + 2 * 72
        # a changed OBJ makes it stale
        obj.write_text(obj.read_text() + "f 2//1 3//1 4//1\n")
        changed = frontend.parse_scene(str(rto))
        assert changed["vertices"].shape[0] == 3 * 6 and cache.stat().st_size == 48 + 3 * 72
        frontend.set_obj_cache(False)
        cache.unlink()
        frontend.parse_scene(str(rto))
        assert not cache.exists()
    finally:
        frontend.set_obj_cache(None)


def test_obj_float_parsing_matches_strtof(frontend, tmp_path):
    """The loader's own decimal -> float conversion (host/parse_obj.cc: to_float, one double operation + midpoint
    check) must give the bits of strtof — what the reference's std::stof returns — for every token: random decimals
    of every shape, exact float midpoints (where rounding through double would differ), range limits, signed zeros."""
    import ctypes
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    libc = ctypes.CDLL(None)
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    rng = np.random.default_rng(123)
    toks = ["0", "-0", "+0.0", "1.", ".5", "-.25", "+3", "1e5", "1E-5", "2.5e+3", "123456789012345", "1234567890123456",
            "0.000000000000000000001", "3.4028234e38", "3.4028236e38", "1.17549435e-38", "1e-45", "7e-46", "1e22", "1e23",
            "16777217", "16777219", "0.1", "0.30000001192092896", "9007199254740993", "1e0005", "00012.5", "1.5000000000000000000"]
    for _ in range(30000):
        nd = int(rng.integers(1, 18))
        digits = "".join(str(int(x)) for x in rng.integers(0, 10, nd))
        dot = int(rng.integers(0, nd + 1))
        t = digits[:dot] + ("." + digits[dot:] if rng.random() < 0.8 else digits[dot:])
        if t in (".", ""):
            t = "0"
        if rng.random() < 0.3:
            t += "eE"[int(rng.integers(0, 2))] + ["", "+", "-"][int(rng.integers(0, 3))] + str(int(rng.integers(0, 40)))
        toks.append(["", "-", "+"][int(rng.integers(0, 3))] + t)
    for _ in range(4000):  # exact midpoints between neighbouring floats, and their neighbourhood, at several lengths
        f = np.float32(rng.standard_normal() * 10.0 ** float(rng.integers(-6, 7)))
        g = np.nextafter(f, np.float32(np.inf))
        mid = (Decimal(float(f)) + Decimal(float(g))) / 2
        toks.append(format(mid, "f"))
        toks.append(format(mid.quantize(Decimal(1).scaleb(mid.adjusted() - int(rng.integers(6, 16)))), "f"))
        toks.append("%.9g" % float(f))
    toks = [t for t in toks if len(t) < 60]
    obj = tmp_path / "f.obj"
    n = len(toks) // 3 * 3
    with open(obj, "w") as fh:
        for i in range(0, n, 3):
            fh.write("v %s %s %s\n" % (toks[i], toks[i + 1], toks[i + 2]))
        fh.write("vn 0 0 1\n")
        for i in range(n // 3 // 3):
            fh.write("f %d//1 %d//1 %d//1\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
    rto = tmp_path / "f.rto"
    rto.write_text("material g { color = (1, 1, 1, 1)\n roughness = 1 }\nmesh { obj_file = %s\n material = g }\n"
                   "camera { position = (0,0,1) look_at = (0,0,0) fov = 40 }\nnum_samples = 1\nnum_bounces = 1\n"
                   "width = 4\nheight = 4\noutput_image = x.ppm\n" % obj)
    got = frontend.parse_scene(str(rto))["vertices"].reshape(-1)
    used = toks[:len(got)]
    want = np.array([libc.strtof(t.encode(), None) for t in used], dtype=np.float32)
    bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0]
    assert len(bad) == 0, [(used[i], float(got[i]), float(want[i])) for i in bad[:10]]
    assert len(got) > 40000


@pytest.mark.parametrize("face,tail", [("f 1// 2//2 3//3", "\n"), ("f 1//1 2//1 3//", ""), ("f 1//1 2//1 3//", "\n"),
                                       ("f 1//1 2//1 -3//1", "\n"), ("f 1//1 2//1 3//1234567890123456789012", "\n"),
                                       ("f 1//1 2//1 3/1", "\n")])
@pytest.mark.parametrize("nommap", ["0", "1"])
def test_obj_malformed_face_indices_are_errors(frontend, tmp_path, monkeypatch, face, tail, nommap):
    """An empty, negative or absurdly long index field is an error confined to its own field: the loader never reads the
    next token (or past the end of the text, with the file ending right behind the field) to find a number.  The
    reference's std::stoi throws on the same input (parse_obj.cc:52-54)."""
    if nommap == "1":
        monkeypatch.setenv("LISA_OBJ_NO_MMAP", "1")   # heap buffer instead of the page-cache mapping
    else:
        monkeypatch.delenv("LISA_OBJ_NO_MMAP", raising=False)
    obj = tmp_path / "bad.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvn 0 1 0\nvn 1 0 0\n" + face + tail)
    rto = tmp_path / "bad.rto"
    rto.write_text("material g { color = (1, 1, 1, 1)\n roughness = 1 }\nmesh { obj_file = %s\n material = g }\n"
                   "camera { position = (0,0,1) look_at = (0,0,0) fov = 40 }\nnum_samples = 1\nnum_bounces = 1\n"
                   "width = 4\nheight = 4\noutput_image = x.ppm\n" % obj)
    with pytest.raises(frontend.SceneError, match="index out of range|without normal index"):
        frontend.parse_scene(str(rto))
