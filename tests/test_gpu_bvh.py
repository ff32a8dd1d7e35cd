"""The GPU BVH builder and both traversal kernels against the oracle's exact closest hit (double
precision, oracle/cpu_ref.c) and against brute-force rules, on the Cornell box and on synthetic soups;
edge cases: empty scene, one triangle, only emitters, duplicated triangles, degenerate (zero-area)
triangles, rays along edges and through vertices of a tessellated sheet (watertightness)."""
import numpy as np
import pytest

from conftest import resized

pytestmark = pytest.mark.gpu

MAT_W = dict(emit=False, alpha=1.0, diffuse=(0.8, 0.8, 0.8), roughness=1.0)
MAT_L = dict(emit=True, alpha=1.0, emission=(1.0, 1.0, 1.0))


def _mk(rt, v, n, m, mats, bvh):
    return rt.Renderer(v, n, m, mats, 8, 8, (0, 0, 5), (0, 0, 0), 45.0, 1, 3, bvh_kind=bvh)


def _rays(rng, lo, hi, n):
    o = rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.uniform(0.25, 3.0, size=(n, 1)).astype(np.float32)  # directions are not unit (Q5)
    return o, d


def _soup(rng, T, size=0.08):
    c = rng.uniform(-1, 1, size=(T, 1, 3))
    v = (c + rng.normal(scale=size, size=(T, 3, 3))).astype(np.float32).reshape(-1, 3)
    n = rng.normal(size=(T * 3, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return v, n


def _compare_with_oracle(R, S, o, d, m=3000):
    prim, t = R.trace_closest(o, d)
    bad = 0
    for i in range(m):
        p, tt = S.closest_hit(o[i], d[i])
        if p != prim[i]:
            if p >= 0 and prim[i] >= 0 and abs(tt - t[i]) <= 2e-5 * max(1.0, abs(tt)):
                continue  # tie between coplanar / touching triangles
            bad += 1
        elif p >= 0:
            assert abs(tt - t[i]) <= 2e-5 * max(1.0, abs(tt))
    return bad, prim


@pytest.mark.parametrize("bvh", [0, 1], ids=["wide8", "binary"])
def test_cornell_closest_and_shadow(rt, orc, cornell, bvh):
    R = rt.Renderer.from_scene(resized(cornell, 16), bvh_kind=bvh)
    st = R.stats()
    assert st["num_triangles"] == 1002 and st["num_emitter_triangles"] == 2
    S = orc.Scene(cornell["vertices"], cornell["normals"], cornell["mat_indices"], cornell["materials_packed"])
    rng = np.random.default_rng(3)
    o, d = _rays(rng, cornell["vertices"].min(0), cornell["vertices"].max(0), 100000)
    bad, prim = _compare_with_oracle(R, S, o, d)
    assert bad <= 2, bad  # float vs double at grazing edges
    assert 0.3 < (prim >= 0).mean() < 0.95
    # shadow rule: outcome decided by the CLOSEST hit (LISA_SHADOW_CLOSEST)
    oc, light = R.trace_shadow(o, d)
    emit = np.array([1 if m["emit"] else 0 for m in cornell["materials"]])
    mat_of = cornell["mat_indices"][np.maximum(prim, 0)]
    exp = np.where(prim < 0, 0, np.where(emit[mat_of] == 1, 1, 2))
    assert (exp != oc).mean() < 1e-4
    assert (light[oc == 1] == 4).all()  # the README scene's light is material 4


@pytest.mark.parametrize("bvh", [0, 1], ids=["wide8", "binary"])
@pytest.mark.parametrize("T", [1, 2, 3, 4, 9, 64, 1000, 50000])
def test_random_soup(rt, orc, bvh, T):
    rng = np.random.default_rng(T)
    v, n = _soup(rng, T)
    m = (rng.random(T) < 0.1).astype(np.int32)  # ~10 % emitters, interleaved
    R = _mk(rt, v, n, m, [MAT_W, MAT_L], bvh)
    assert R.stats()["num_emitter_triangles"] == int(m.sum())
    S = orc.Scene(v, n, m, rt_pack([MAT_W, MAT_L]))
    o, d = _rays(rng, v.min(0), v.max(0), 20000)
    bad, prim = _compare_with_oracle(R, S, o, d, m=1500)
    assert bad <= 1
    oc, light = R.trace_shadow(o, d)
    exp = np.where(prim < 0, 0, np.where(m[np.maximum(prim, 0)] == 1, 1, 2))
    assert (exp != oc).mean() < 2e-4


@pytest.mark.parametrize("scale,offset", [(1.0, (0, 0, 0)), (1e-3, (0, 0, 0)), (1e3, (0, 0, 0)), (1.0, (300.0, -200.0, 500.0)),
                                          (1e3, (4e4, 1e4, -3e4))], ids=["unit", "milli", "kilo", "far", "kilo_far"])
def test_node_planes_conservative_at_any_scale(rt, orc, scale, offset):
    """The compressed nodes' planes are evaluated as fma(1 + q 2^-15, 2^15 a, o - 2^15 a) with an absolute pad
    (traverse.cuh): the pad must cover the cancellation for small and large scenes and for scenes far from the origin
    (|origin term| >> node extent), or hits would be lost.  Axis-parallel rays (reciprocal clamped) included."""
    rng = np.random.default_rng(11)
    T = 20000
    v, n = _soup(rng, T, size=0.03)
    v = (v * np.float32(scale) + np.float32(offset)).astype(np.float32)
    m = (rng.random(T) < 0.05).astype(np.int32)
    R = _mk(rt, v, n, m, [MAT_W, MAT_L], 0)
    S = orc.Scene(v, n, m, rt_pack([MAT_W, MAT_L]))
    o, d = _rays(rng, v.min(0), v.max(0), 6000)
    d[:1000, rng.integers(0, 3, 1000)[0]] = 0.0          # one component exactly zero
    d[1000:1500] = np.float32([0, 0, -1]) * np.float32(scale)  # axis-parallel
    prim, t = R.trace_closest(o, d)
    bad = 0
    for i in range(2500):
        p, tt = S.closest_hit(o[i], d[i])
        if p != prim[i] and not (p >= 0 and prim[i] >= 0 and abs(tt - t[i]) <= 2e-5 * max(1.0, abs(tt))):
            bad += 1
    # float vs double: far from the origin the triangle test itself loses bits at grazing edges, the boxes must not add to it
    assert bad <= (2 if offset == (0, 0, 0) else 12), bad
    assert 0.05 < (prim >= 0).mean() < 0.999


def test_chunked_upload_is_exact(rt, monkeypatch):
    """lisa_create's pinned, chunked upload (devmem.cu: upload_async, used above 1 GB per array) forced on a 1.5M-triangle
    soup (54 MB of vertices = 4 chunks over the 4 workers): same BVH and same hits as the plain copy."""
    rng = np.random.default_rng(5)
    T = 1_500_000
    c = rng.random((T, 1, 3), dtype=np.float32)
    v = (c + (rng.random((T, 3, 3), dtype=np.float32) - 0.5) * np.float32(0.01)).reshape(-1, 3)
    n = np.repeat(np.float32([[0, 1, 0]]), 3 * T, axis=0)
    m = np.zeros(T, np.int32); m[-2:] = 1
    o, d = _rays(rng, v.min(0), v.max(0), 50000)
    res = []
    for thresh in ("0", "1000000000000"):
        monkeypatch.setenv("LISA_UPLOAD_CHUNKED_MIN", thresh)
        R = _mk(rt, v, n, m, [MAT_W, MAT_L], 0)
        res.append(R.trace_closest(o, d) + (R.stats()["bvh_nodes"],))
        R.close()
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1], res[1][1])
    assert res[0][2] == res[1][2]              # the builder is deterministic: same data, same tree
    assert (res[0][0] >= 0).mean() > 0.05


def rt_pack(mats):
    from oracle.scene_py import pack_material
    return b"".join(pack_material(roughness=m.get("roughness", 0), alpha=m["alpha"], n=m.get("n", 0),
                                  diffuse=m.get("diffuse", (0, 0, 0)), emit=m["emit"], emission=m.get("emission", (0, 0, 0)))
                    for m in mats)


@pytest.mark.parametrize("bvh", [0, 1], ids=["wide8", "binary"])
def test_empty_and_degenerate_scenes(rt, bvh):
    z3 = np.zeros((0, 3), np.float32)
    R = _mk(rt, z3, z3, np.zeros(0, np.int32), [MAT_W], bvh)
    o = np.zeros((4, 3), np.float32)
    d = np.ones((4, 3), np.float32)
    prim, _ = R.trace_closest(o, d)
    assert (prim == -1).all()
    R.render_subframes(0, 1, 2)
    assert float(np.abs(R.read_accum()[..., :3]).max()) == 0.0  # every ray misses: black (bg = 0)
    # only emitters; only one triangle; duplicates; zero-area triangles
    tri = np.float32([[-1, -1, 0], [1, -1, 0], [0, 1, 0]])
    nrm = np.float32([[0, 0, 1]] * 3)
    R = _mk(rt, tri, nrm, np.zeros(1, np.int32), [MAT_L], bvh)
    prim, t = R.trace_closest(np.float32([[0, 0, 2]]), np.float32([[0, 0, -1]]))
    assert prim[0] == 0 and abs(t[0] - 2) < 1e-6
    oc, light = R.trace_shadow(np.float32([[0, 0, 2], [5, 5, 2]]), np.float32([[0, 0, -1], [0, 0, -1]]))
    assert oc.tolist() == [1, 0] and light[0] == 0
    dup = np.concatenate([tri] * 40 + [np.float32([[3, 3, 3]] * 3)] * 5)  # 40 identical + 5 zero-area triangles
    R = _mk(rt, dup, np.concatenate([nrm] * 45), np.zeros(45, np.int32), [MAT_W], bvh)
    prim, t = R.trace_closest(np.float32([[0, 0, 2], [0.2, -0.5, -3]]), np.float32([[0, 0, -0.5], [0, 0, 2]]))
    assert (prim >= 0).all() and (prim < 40).all()
    np.testing.assert_allclose(t, [4.0, 1.5], rtol=1e-6)  # t is in units of |dir|


@pytest.mark.parametrize("bvh", [0, 1], ids=["wide8", "binary"])
def test_watertight_sheet(rt, bvh):
    """A 32x32 tessellated, slightly warped sheet: rays through shared edges and vertices must not leak."""
    k = 32
    rng = np.random.default_rng(5)
    gx, gy = np.meshgrid(np.linspace(-1, 1, k + 1), np.linspace(-1, 1, k + 1), indexing="ij")
    gz = 0.05 * np.sin(3 * gx) * np.cos(2 * gy)
    P = np.stack([gx, gy, gz], -1).astype(np.float32)
    tris = []
    for i in range(k):
        for j in range(k):
            a, b, c, d = P[i, j], P[i + 1, j], P[i + 1, j + 1], P[i, j + 1]
            tris += [a, b, c, a, c, d]
    v = np.array(tris, np.float32)
    n = np.tile(np.float32([[0, 0, 1]]), (v.shape[0], 1))
    R = _mk(rt, v, n, np.zeros(v.shape[0] // 3, np.int32), [MAT_W], bvh)
    # targets ON vertices, ON edge midpoints and along edges, from random origins above the sheet
    inner = P[1:-1, 1:-1].reshape(-1, 3)
    mids = 0.5 * (P[1:-1, 1:-1] + P[2:, 1:-1]).reshape(-1, 3)
    diag = 0.5 * (P[:-1, :-1] + P[1:, 1:]).reshape(-1, 3)
    lam = rng.random((inner.shape[0], 1)).astype(np.float32)
    along = (P[1:-1, 1:-1].reshape(-1, 3) * lam + P[1:-1, 2:].reshape(-1, 3) * (1 - lam))
    tgt = np.concatenate([inner, mids, diag, along]).astype(np.float32)
    tgt = np.concatenate([tgt] * 8)
    o = (tgt + np.concatenate([rng.normal(scale=0.5, size=(tgt.shape[0], 2)), rng.uniform(0.5, 3, size=(tgt.shape[0], 1))], 1)).astype(np.float32)
    prim, t = R.trace_closest(o, (tgt - o).astype(np.float32), tmin=0.0)
    assert (prim >= 0).all(), "%d rays leaked through the sheet" % (prim < 0).sum()
    oc, _ = R.trace_shadow(o, (tgt - o).astype(np.float32), tmin=0.0)
    assert (oc == 2).all()


def test_wide_and_binary_agree(rt, cornell):
    rng = np.random.default_rng(11)
    o, d = _rays(rng, cornell["vertices"].min(0), cornell["vertices"].max(0), 200000)
    res = []
    for bvh in (0, 1):
        R = rt.Renderer.from_scene(resized(cornell, 16), bvh_kind=bvh)
        res.append(R.trace_closest(o, d))
    same = res[0][0] == res[1][0]
    assert same.mean() > 0.9999  # only exact ties may resolve differently
    np.testing.assert_allclose(res[0][1][same], res[1][1][same], rtol=1e-6)


def test_tmin_tmax_in_parametric_units(rt):
    tri = np.float32([[-1, -1, 0], [1, -1, 0], [0, 1, 0]])
    nrm = np.float32([[0, 0, 1]] * 3)
    R = _mk(rt, tri, nrm, np.zeros(1, np.int32), [MAT_W], 0)
    o = np.float32([[0, 0, 1]] * 3)
    d = np.float32([[0, 0, -4], [0, 0, -4], [0, 0, -4]])  # hit at t = 0.25
    assert R.trace_closest(o, d, tmin=1e-4, tmax=1e16)[0][0] == 0
    assert R.trace_closest(o, d, tmin=0.3, tmax=1e16)[0][0] == -1
    assert R.trace_closest(o, d, tmin=1e-4, tmax=0.2)[0][0] == -1


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 256, 2047, 2048, 2049, 4097, 100003, 3_000_001])
def test_builder_primitives(rt, n):
    """Own LSD radix sort (stable, 64-bit keys), exclusive scan and compaction against numpy; bit-exact."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    if n > 100:
        keys[rng.integers(0, n, n // 3)] = keys[0]          # many duplicates: stability matters
        keys[rng.integers(0, n, n // 7)] |= np.uint64(1) << np.uint64(63)
    vals = np.arange(n, dtype=np.uint32)
    k, v = rt.debug_sort_pairs(keys, vals)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(k, keys[order])
    np.testing.assert_array_equal(v, vals[order])
    a = rng.integers(0, 1000, size=n, dtype=np.uint32)
    c = rng.integers(-3, 50, size=n).astype(np.int32)
    s, tot, kept = rt.debug_scan_compact(a, c)
    np.testing.assert_array_equal(s, (np.cumsum(a, dtype=np.uint64) - a).astype(np.uint32))
    assert tot == int(a.sum(dtype=np.uint64)) & 0xffffffff
    np.testing.assert_array_equal(kept, c[c >= 0])


@pytest.mark.parametrize("pipe", ["pool", "path", "wavefront"])
def test_every_schedule_on_empty_and_emitter_only_scenes(rt, pipe, monkeypatch):
    """No non-emitter triangle at all (root_other < 0): a ray that misses the emitter bounds has nothing to traverse and
    must come back as a MISS.  k_pool used to park such a slot with the ray set-up in the hit record and read a bogus
    primitive; small images go to k_path by default, so every schedule is forced here."""
    monkeypatch.setenv("LISA_PIPELINE", pipe)
    z3 = np.zeros((0, 3), np.float32)
    R = rt.Renderer(z3, z3, np.zeros(0, np.int32), [MAT_W], 64, 48, (0, 0, 5), (0, 0, 0), 45.0, 1, 3)
    R.render_subframes(0, 2, 3)
    assert float(np.abs(R.read_accum()[..., :3]).max()) == 0.0
    R.close()
    # one small emitter triangle in the middle of the view: most camera rays miss its bounds, the others see it
    tri = np.float32([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0, 0.5, 0]])
    nrm = np.float32([[0, 0, 1]] * 3)
    R = rt.Renderer(tri, nrm, np.zeros(1, np.int32), [MAT_L], 64, 48, (0, 0, 5), (0, 0, 0), 45.0, 1, 3)
    R.render_subframes(0, 2, 3)
    a = R.read_accum()[..., :3]
    st = R.stats()
    assert np.isfinite(a).all() and a.max() == 1.0 and a.min() == 0.0
    assert 0.005 < (a[..., 0] > 0).mean() < 0.2           # the triangle covers a small part of the image
    assert a[24, 32, 0] == 1.0 and a[0, 0, 0] == 0.0      # centre pixel sees the emitter (emission 1), corner misses
    assert st["last_samples"] == 64 * 48 * 6 and st["last_radiance_rays"] == st["last_samples"]
    assert st["last_shadow_rays"] == 0                    # no opaque hit, no light sampling
    R.close()


@pytest.mark.parametrize("bvh", [0, 1], ids=["wide8", "binary"])
def test_ten_thousand_coincident_triangles(rt, bvh):
    """10,000 copies of one triangle (identical Morton keys, identical boxes, every merged area a tie) plus a floor: the
    PLOC tie-break pairs such runs up, so the hierarchy stays balanced — it builds in ~14 rounds instead of ~10,000 and
    no ray overflows its traversal stack.  Every ray through the stack hits one of the copies at the exact distance."""
    tri = np.float32([[-1, -1, 0], [1, -1, 0], [0, 1, 0]])
    floor = np.float32([[-5, -5, -1], [5, -5, -1], [0, 5, -1]])
    T = 10000
    v = np.concatenate([np.tile(tri, (T, 1)), floor])
    n = np.tile(np.float32([[0, 0, 1]]), (3 * (T + 1), 1))
    m = np.zeros(T + 1, np.int32)
    R = _mk(rt, v, n, m, [MAT_W], bvh)
    o = np.float32([[0, 0, 2], [0.1, -0.3, 3], [3, -3, 2], [9, 9, 2]])
    d = np.float32([[0, 0, -1], [0, 0, -2], [0, 0, -1], [0, 0, -1]])
    prim, t = R.trace_closest(o, d)
    assert 0 <= prim[0] < T and 0 <= prim[1] < T and prim[2] == T and prim[3] == -1
    np.testing.assert_allclose(t[:3], [2.0, 1.5, 3.0], rtol=1e-6)
    R.render_subframes(0, 1, 1)                           # the render kernels traverse it too
    assert np.isfinite(R.read_accum()).all()


def test_traversal_stack_overflow_is_loud(built, tmp_path):
    """A traversal stack that fills up drops a subtree; the call that ran the kernel must then FAIL (LISA_ERR_STATE) instead
    of returning hits that may be wrong.  The builders do not bound the depth of the tree, so the condition is provoked
    with a test build of the same sources whose stacks hold 3 entries (make liblisa_rt_tinystack.so): a 50k-triangle soup
    overflows them on most rays, in the diagnostic queries and in the render kernels alike."""
    import os, subprocess, sys
    from conftest import ROOT
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lisa_b200"), "liblisa_rt_tinystack.so"])
    code = ("import numpy as np, lisa_b200.rt as rt\n"
            "rng = np.random.default_rng(1); T = 50000\n"
            "c = rng.uniform(-1, 1, size=(T, 1, 3)); v = (c + rng.normal(scale=0.3, size=(T, 3, 3))).astype(np.float32).reshape(-1, 3)\n"
            "n = np.tile(np.float32([[0, 0, 1]]), (3 * T, 1)); m = np.zeros(T, np.int32)\n"
            "mats = [dict(emit=False, alpha=1.0, diffuse=(0.8, 0.8, 0.8), roughness=1.0)]\n"
            "R = rt.Renderer(v, n, m, mats, 32, 32, (0, 0, 5), (0, 0, 0), 45.0, 1, 3)\n"
            "o = rng.uniform(-2, 2, size=(4096, 3)).astype(np.float32); d = rng.normal(size=(4096, 3)).astype(np.float32)\n"
            "out = []\n"
            "for call in (lambda: R.trace_closest(o, d), lambda: R.trace_shadow(o, d), lambda: R.render_subframes(0, 1, 2)):\n"
            "    try:\n"
            "        call(); out.append('ok')\n"
            "    except rt.LisaError as e:\n"
            "        out.append('%d %s' % (e.code, e))\n"
            "print('\\n'.join(out))\n")
    for lib, expect_fail in (("liblisa_rt_tinystack.so", True), ("liblisa_rt.so", False)):
        env = dict(os.environ, LISA_RT_LIB=os.path.join(ROOT, "lisa_b200", lib))
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = r.stdout.strip().splitlines()
        assert len(lines) == 3
        for ln in lines:
            if expect_fail:
                assert ln.startswith("-5 ") and "traversal stack overflow" in ln, ln
            else:
                assert ln == "ok", ln


def test_serialised_bvh_round_trip(rt, cornell, tmp_path):
    """lisa_save_bvh / lisa_create_from_bvh: a context made from the file — no geometry handed over,
    no build — answers closest-hit queries and renders bit-identically to the context that wrote it; truncated files, files
    for another triangle count and files built for other emitter flags are refused."""
    sc = resized(cornell, 96)
    A = rt.Renderer.from_scene(sc)
    path = str(tmp_path / "cornell.lisabvh")
    A.save_bvh(path)
    stA = A.stats()
    import os
    assert os.path.getsize(path) == 136 + stA["bvh_bytes"] + stA["num_triangles"] * 100   # header (version 3), nodes, 48 + 48 + 4 bytes per triangle
    B = rt.Renderer.from_bvh(sc, path)
    stB = B.stats()
    for k in ("num_triangles", "num_emitter_triangles", "bvh_nodes", "bvh_emitter_nodes", "bvh_bytes", "bvh_sah_nodes_per_ray", "pool_flavour"):
        assert stA[k] == stB[k], k
    assert stA["bvh_sah_nodes_per_ray"] > 1.0   # the estimate that chooses k_pool's flavour travels with the file
    assert stB["bvh_build_ms"] == 0.0
    rng = np.random.default_rng(9)
    o, d = _rays(rng, cornell["vertices"].min(0), cornell["vertices"].max(0), 50000)
    pa, ta = A.trace_closest(o, d)
    pb, tb = B.trace_closest(o, d)
    np.testing.assert_array_equal(pa, pb)
    np.testing.assert_array_equal(ta, tb)
    A.render_subframes(0, 2, 8)
    B.render_subframes(0, 2, 8)
    np.testing.assert_array_equal(A.read_accum(), B.read_accum())
    C = rt.Renderer.from_bvh(sc, path, with_geometry=True)   # geometry present: only its triangle count is checked
    C.render_subframes(0, 2, 8)
    np.testing.assert_array_equal(A.read_accum(), C.read_accum())
    for R in (A, B, C):
        R.close()
    raw = open(path, "rb").read()
    bad = str(tmp_path / "cut.lisabvh")
    open(bad, "wb").write(raw[:-1000])
    with pytest.raises(rt.LisaError, match="truncated or damaged"):
        rt.Renderer.from_bvh(sc, bad)
    open(bad, "wb").write(b"P6\n" + raw[3:])
    with pytest.raises(rt.LisaError, match="not a serialised BVH"):
        rt.Renderer.from_bvh(sc, bad)
    other = dict(sc)
    other["vertices"], other["normals"], other["mat_indices"] = sc["vertices"][:-3], sc["normals"][:-3], sc["mat_indices"][:-1]
    with pytest.raises(rt.LisaError, match="holds 1002 triangles, the scene has 1001"):
        rt.Renderer.from_bvh(other, path, with_geometry=True)
    swapped = dict(sc)
    swapped.pop("materials_packed", None)
    swapped["materials"] = [dict(m) for m in sc["materials"]]
    swapped["materials"][0], swapped["materials"][4] = swapped["materials"][4], swapped["materials"][0]   # the light becomes material 0
    with pytest.raises(rt.LisaError, match="other emitter flags"):
        rt.Renderer.from_bvh(swapped, path)


def _needle_soup(rng, T_small, T_long):
    """Small triangles plus long thin diagonal ones, whose bounding boxes are almost entirely empty."""
    v, n = _soup(rng, T_small, size=0.02)
    a = rng.uniform(-1, 1, size=(T_long, 3))
    b = a + rng.uniform(0.8, 1.6, size=(T_long, 1)) * rng.choice([-1.0, 1.0], size=(T_long, 3))
    w = rng.normal(scale=0.01, size=(T_long, 3))
    vl = np.stack([a, b, a + w], axis=1).astype(np.float32).reshape(-1, 3)
    nl = rng.normal(size=(T_long * 3, 3)).astype(np.float32)
    nl /= np.linalg.norm(nl, axis=1, keepdims=True)
    return np.concatenate([v, vl]), np.concatenate([n, nl])


@pytest.mark.parametrize("bvh", [0, 1], ids=["wide8", "binary"])
def test_split_triangles_same_hits_less_work(rt, orc, bvh):
    """LISA_FLAG_SPLIT_TRIANGLES (csrc/split.cu): the BVH is built over references with clipped boxes; every ray finds the
    hit it finds in the unsplit tree and the oracle's, the rendered image is the same bits, the traversal does less work."""
    rng = np.random.default_rng(77)
    T_small, T_long = 30000, 1500
    v, n = _needle_soup(rng, T_small, T_long)
    T = T_small + T_long
    m = np.zeros(T, np.int32)
    m[rng.choice(T_small, 300, replace=False)] = 1
    m[T_small:T_small + 20] = 1  # some long emitters too: the emitter partition is split as well
    args = (v, n, m, [MAT_W, MAT_L], 96, 96, (0, 0, 4.5), (0, 0, 0), 40.0, 4, 4)
    R0 = rt.Renderer(*args, bvh_kind=bvh)
    R1 = rt.Renderer(*args, bvh_kind=bvh, flags=rt.FLAG_SPLIT_TRIANGLES)
    s0, s1 = R0.stats(), R1.stats()
    assert s0["num_references"] == s0["num_triangles"] == T
    assert s1["num_triangles"] == T and T + T_long < s1["num_references"] <= 1.5 * T + 64
    assert s1["num_emitter_triangles"] > s0["num_emitter_triangles"] == 320   # references of the emitter partition
    o, d = _rays(rng, v.min(0), v.max(0), 200000)
    p0, t0 = R0.trace_closest(o, d)
    p1, t1 = R1.trace_closest(o, d)
    same = (p0 == p1)
    assert (t0[same] == t1[same]).all()
    assert (~same).sum() <= 2 and np.allclose(t0[~same], t1[~same], rtol=2e-5)  # ties only
    assert (p1 >= T_small).mean() > 0.02   # the long triangles are hit
    S = orc.Scene(v, n, m, rt_pack([MAT_W, MAT_L]))
    bad, _ = _compare_with_oracle(R1, S, o, d, m=2000)
    assert bad <= 2, bad
    oc0, l0 = R0.trace_shadow(o, d)
    oc1, l1 = R1.trace_shadow(o, d)
    assert (oc0 != oc1).sum() <= 2
    R0.render(); R1.render()
    assert np.array_equal(R0.read_accum(), R1.read_accum())
    w0, w1 = R0.stats(), R1.stats()
    assert w1["triangles_tested"] < 0.8 * w0["triangles_tested"], (w0["triangles_tested"], w1["triangles_tested"])
    assert w1["nodes_visited"] < w0["nodes_visited"], (w0["nodes_visited"], w1["nodes_visited"])


def test_split_triangles_leaves_short_triangles_alone(rt, cornell, tmp_path):
    """A soup without long triangles gets one reference per triangle; degenerate and empty soups go through; a split BVH
    survives lisa_save_bvh / lisa_create_from_bvh."""
    rng = np.random.default_rng(5)
    v, n = _soup(rng, 20000, size=0.01)
    m = np.zeros(20000, np.int32)
    R = rt.Renderer(v, n, m, [MAT_W], 8, 8, (0, 0, 5), (0, 0, 0), 45.0, 1, 3, flags=rt.FLAG_SPLIT_TRIANGLES)
    assert R.stats()["num_references"] == 20000
    # zero-area triangles and one huge triangle
    v2 = np.zeros((9, 3), np.float32)
    v2[3:6] = [[-5, -5, 0], [5, -5, 0], [0, 7, 0]]
    v2[6:9] = [[1, 1, 1], [2, 2, 2], [3, 3, 3]]
    R2 = rt.Renderer(v2, np.ones((9, 3), np.float32), np.zeros(3, np.int32), [MAT_W], 8, 8, (0, 0, 5), (0, 0, 0), 45.0, 1, 3,
                     flags=rt.FLAG_SPLIT_TRIANGLES)
    st = R2.stats()
    assert st["num_triangles"] == 3 and 3 <= st["num_references"] <= 3 * 64
    prim, t = R2.trace_closest(np.array([[0.3, 0.2, 3.0], [40.0, 0, 3.0]], np.float32), np.array([[0, 0, -1.0], [0, 0, -1.0]], np.float32))
    assert prim[0] == 1 and abs(t[0] - 3.0) < 1e-5 and prim[1] == -1
    R3 = rt.Renderer(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.int32), [MAT_W], 8, 8, (0, 0, 5),
                     (0, 0, 0), 45.0, 1, 3, flags=rt.FLAG_SPLIT_TRIANGLES)
    assert R3.stats()["num_references"] == 0
    # Cornell: the walls are axis-aligned right triangles, whose boxes are as full as a triangle's box gets -> left alone
    sc = resized(cornell, 48)
    A = rt.Renderer.from_scene(sc, flags=rt.FLAG_SPLIT_TRIANGLES)
    assert A.stats()["num_references"] == A.stats()["num_triangles"] == 1002
    C = rt.Renderer.from_scene(sc)
    A.render(); C.render()
    assert np.array_equal(A.read_accum(), C.read_accum())
    # a split BVH through lisa_save_bvh / lisa_create_from_bvh: references and images survive
    v, n = _needle_soup(rng, 5000, 300)
    m = np.zeros(5300, np.int32); m[:40] = 1
    sc = dict(vertices=v, normals=n, mat_indices=m, materials=[MAT_W, MAT_L], width=64, height=64,
              camera=dict(eye=(0, 0, 4.5), look_at=(0, 0, 0), fov=40.0), num_samples=4, num_bounces=4)
    A = rt.Renderer.from_scene(sc, flags=rt.FLAG_SPLIT_TRIANGLES)
    sa = A.stats()
    assert sa["num_triangles"] == 5300 and sa["num_references"] > 5300 + 300
    path = str(tmp_path / "split.bvh")
    A.save_bvh(path)
    B = rt.Renderer.from_bvh(sc, path)
    sb = B.stats()
    assert sb["num_triangles"] == 5300 and sb["num_references"] == sa["num_references"]
    with pytest.raises(rt.LisaError):
        sc2 = dict(sc, vertices=v[:-3], normals=n[:-3], mat_indices=m[:-1])
        rt.Renderer.from_bvh(sc2, path, with_geometry=True)   # 5299 triangles against a file built from 5300
    A.render(); B.render()
    assert np.array_equal(A.read_accum(), B.read_accum())


def _overlap_soup(rng, T):
    """triangles about as large as their spacing, like the C4 soups: neighbours in Morton order overlap but share no box"""
    edge = 0.5 * T ** (-1.0 / 3.0)
    c = rng.random((T, 1, 3))
    v = (c + (rng.random((T, 3, 3)) - 0.5) * 2 * edge).astype(np.float32).reshape(-1, 3)
    q = np.float32([[-0.5, 1.6, -0.5], [1.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, -0.5], [1.5, 1.6, 1.5], [-0.5, 1.6, 1.5]])
    v = np.concatenate([v, q])
    n = np.tile(np.float32([[0, 1, 0]]), (len(v), 1))
    m = np.concatenate([np.zeros(T, np.int32), np.ones(2, np.int32)])
    return v, n, m


def test_sah_leaves_same_hits_less_work(rt, orc, cornell, monkeypatch):
    """Leaves by the surface-area heuristic (k_collapse8, LISA_LEAF_SAH; default on): a subtree of 2-3 triangles becomes ONE leaf
    only where that costs no more triangle tests than keeping its halves apart.  On a soup of overlapping triangles the
    traversal tests far fewer triangles and finds the same hits (those of the oracle); on a mesh whose neighbours share
    their boxes (the Cornell box: quads, a tessellated sphere) the tree stays what it was."""
    rng = np.random.default_rng(5)
    T = 60000
    v, n, m = _overlap_soup(rng, T)
    args = (v, n, m, [MAT_W, MAT_L], 96, 96, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, 4, 4)
    R1 = rt.Renderer(*args)
    monkeypatch.setenv("LISA_LEAF_SAH", "-1")   # every subtree of <= 3 triangles is one leaf
    R0 = rt.Renderer(*args)
    monkeypatch.delenv("LISA_LEAF_SAH")
    assert R1.stats()["bvh_nodes"] > R0.stats()["bvh_nodes"]
    o, d = _rays(rng, v.min(0), v.max(0), 200000)
    p0, t0 = R0.trace_closest(o, d)
    p1, t1 = R1.trace_closest(o, d)
    same = (p0 == p1)
    assert (t0[same] == t1[same]).all()
    assert (~same).sum() <= 4 and np.allclose(t0[~same], t1[~same], rtol=2e-5)  # ties between crossing triangles only
    S = orc.Scene(v, n, m, rt_pack([MAT_W, MAT_L]))
    bad, _ = _compare_with_oracle(R1, S, o, d, m=2000)
    assert bad <= 2, bad
    R0.render(); R1.render()
    a0, a1 = R0.read_accum(), R1.read_accum()
    assert (np.abs(a0 - a1).max(axis=2) > 0).mean() < 1e-3   # a tie resolves by test order; every other pixel is the same bits
    w0, w1 = R0.stats(), R1.stats()
    assert w1["triangles_tested"] < 0.75 * w0["triangles_tested"], (w0["triangles_tested"], w1["triangles_tested"])
    assert w1["nodes_visited"] < 1.1 * w0["nodes_visited"], (w0["nodes_visited"], w1["nodes_visited"])
    # the Cornell box: the same tree either way, up to a handful of leaves (152-169 wide nodes, depending on the PLOC radius)
    C1 = rt.Renderer.from_scene(resized(cornell, 32))
    monkeypatch.setenv("LISA_LEAF_SAH", "-1")
    C0 = rt.Renderer.from_scene(resized(cornell, 32))
    assert abs(C1.stats()["bvh_nodes"] - C0.stats()["bvh_nodes"]) <= 0.1 * C0.stats()["bvh_nodes"]


def test_pool_flavours_are_bit_identical_and_chosen_by_the_sah_estimate(rt, cornell, monkeypatch):
    """k_pool's two flavours (sched_pool.cuh: 64 chains per warp and a 4-entry shared-memory stack, or 48 and 12) render the
    same bits; the library picks the deep one where the builder's surface-area estimate says rays visit many nodes."""
    rng = np.random.default_rng(6)
    v, n, m = _overlap_soup(rng, 40000)
    args = (v, n, m, [MAT_W, MAT_L], 160, 160, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, 2, 5)
    monkeypatch.setenv("LISA_PIPELINE", "pool")
    imgs, stats = {}, {}
    for fl in ("auto", "deep", "shallow"):
        if fl == "auto": monkeypatch.delenv("LISA_POOL_FLAVOUR", raising=False)
        else: monkeypatch.setenv("LISA_POOL_FLAVOUR", fl)
        for name, mk in (("soup", lambda: rt.Renderer(*args)), ("cornell", lambda: rt.Renderer.from_scene(resized(cornell, 160)))):
            R = mk()
            stats[name, fl, "created"] = R.stats()
            R.render_subframes(0, 2, 3)
            imgs[name, fl], stats[name, fl] = R.read_accum(), R.stats()
    for name in ("soup", "cornell"):
        assert np.array_equal(imgs[name, "deep"], imgs[name, "shallow"]) and np.array_equal(imgs[name, "auto"], imgs[name, "deep"])
        for k in ("radiance_rays", "shadow_rays", "nodes_visited", "triangles_tested", "shadow_culled"):
            assert stats[name, "deep"][k] == stats[name, "shallow"][k], (name, k)
        assert stats[name, "deep"]["pool_flavour"] == 1 and stats[name, "shallow"]["pool_flavour"] == 0   # forced: stays
    assert 20 < stats["soup", "auto", "created"]["bvh_sah_nodes_per_ray"] < 40 and stats["soup", "auto", "created"]["pool_flavour"] == 0
    assert stats["cornell", "auto"]["bvh_sah_nodes_per_ray"] < 6 and stats["cornell", "auto"]["pool_flavour"] == 0
    monkeypatch.delenv("LISA_POOL_FLAVOUR", raising=False)
    v, n, m = _overlap_soup(rng, 1500000)   # deep from an estimate of 40 (measured crossover: lisa_rt.cu)
    big = (v, n, m, [MAT_W, MAT_L], 128, 128, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, 2, 7)
    st = rt.Renderer(*big).stats()
    assert st["bvh_sah_nodes_per_ray"] > 60 and st["pool_flavour"] == 1
    # ... and whatever the estimate said, the flavour follows the node visits per ray a call has measured
    monkeypatch.setenv("LISA_POOL_DEEP_SAH", "1e9")
    R = rt.Renderer(*big)
    assert R.stats()["pool_flavour"] == 0
    R.render_subframes(0, 1, 2)
    st = R.stats()
    assert st["last_nodes_visited"] / (st["last_radiance_rays"] + st["last_shadow_rays"] - st["last_shadow_culled"]) > 10 and st["pool_flavour"] == 1
