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
