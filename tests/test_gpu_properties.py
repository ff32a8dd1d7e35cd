"""Size-independent properties of the CUDA path at (and around) BASELINE.json's full sizes, plus the
output stage (tonemap, PPM) and the host render()/display() drivers."""
import os

import numpy as np
import pytest

from conftest import ROOT, resized

pytestmark = pytest.mark.gpu


def test_deterministic_and_tiling_invariant(rt, cornell):
    """Chains own their RNG streams: re-running, or running with a different number of resident chains
    (different tiles / queue orders), gives bit-identical accumulators."""
    sc = resized(cornell, 96)
    imgs = []
    for mc in (0, 0, 3000, 96 * 96, 2 * 96 * 96 + 5):
        R = rt.Renderer.from_scene(sc, max_chains=mc)
        R.render_subframes(0, 3, 4)
        imgs.append(R.read_accum())
    for im in imgs[1:]:
        np.testing.assert_array_equal(imgs[0], im)


def test_culling_is_exact(rt, cornell):
    """Shadow tries that cannot change RayState::hit are resolved without traversal (LISA_FLAG_NO_CULL turns
    that off): same RNG consumption, same counts of reference-semantic rays, bit-identical image."""
    sc = resized(cornell, 128)
    out = []
    for flags in (0, rt.FLAG_NO_CULL):
        for bvh in (0, 1):
            R = rt.Renderer.from_scene(sc, flags=flags, bvh_kind=bvh)
            R.render_subframes(0, 2, 6)
            out.append((R.read_accum(), R.stats()))
    base_img, base_st = out[0]
    for img, st in out[1:]:
        np.testing.assert_array_equal(base_img, img)
        assert st["last_shadow_rays"] == base_st["last_shadow_rays"]
        assert st["last_radiance_rays"] == base_st["last_radiance_rays"]
    assert out[0][1]["last_shadow_culled"] > 0.8 * out[0][1]["last_shadow_rays"]
    assert out[2][1]["last_shadow_culled"] == 0


def test_pipelines_are_bit_identical(rt, cornell, monkeypatch):
    """Three schedules of the same estimator: k_path (a lane owns a chain, state in registers), k_pool (a warp owns 64
    chains in shared memory, ready/pending queues) and the wavefront pipeline (three kernels per bounce over chain state
    in HBM).  Every multiply-add is spelled out and the kernels are compiled with -fmad=false, so they run the same
    per-chain arithmetic in the same order: identical accumulators and identical ray / node / triangle counters, for
    every BVH kind, shadow policy and with culling on or off."""
    sc = resized(cornell, 128)
    keys = ("last_radiance_rays", "last_shadow_rays", "last_shadow_culled", "last_shadow_jobs", "null_directions",
            "last_nodes_visited", "last_triangles_tested")
    first = None
    for kw in (dict(), dict(bvh_kind=1), dict(shadow_mode=1), dict(flags=rt.FLAG_NO_CULL), dict(bvh_kind=1, shadow_mode=1)):
        res = {}
        for pipe in ("path", "pool", "wavefront"):
            monkeypatch.setenv("LISA_PIPELINE", pipe)
            R = rt.Renderer.from_scene(sc, **kw)
            R.render_subframes(0, 2, 8)
            res[pipe] = (R.read_accum(), R.stats())
            R.close()
        monkeypatch.delenv("LISA_PIPELINE")
        for pipe in ("pool", "wavefront"):
            np.testing.assert_array_equal(res["path"][0], res[pipe][0])
            for key in keys:
                assert res["path"][1][key] == res[pipe][1][key], (kw, pipe, key)
        assert res["path"][1]["last_kernel_launches"] == res["pool"][1]["last_kernel_launches"] < res["wavefront"][1]["last_kernel_launches"]
        first = first if first is not None else res["path"][0]
    # the option flag selects the wavefront pipeline too
    R = rt.Renderer.from_scene(sc, flags=rt.FLAG_WAVEFRONT)
    R.render_subframes(0, 2, 8)
    np.testing.assert_array_equal(R.read_accum(), first)
    assert R.stats()["last_kernel_launches"] > 10
    R.close()


def test_default_pipeline_choice_is_invisible(rt, cornell, monkeypatch):
    """By default a tile with enough chains to fill k_pool's slots twice runs k_pool, a smaller one k_path: a 640x640
    subframe (409,600 chains) is on the k_pool side of the switch and must equal k_path's image; so must a ragged
    image whose chain count is not a multiple of anything."""
    for w, h in ((640, 640), (333, 17)):
        sc = resized(cornell, w)
        sc["height"] = h
        imgs = []
        for pipe in (None, "path", "pool"):
            if pipe:
                monkeypatch.setenv("LISA_PIPELINE", pipe)
            R = rt.Renderer.from_scene(sc)
            R.render_subframes(3, 1, 2)
            imgs.append(R.read_accum())
            R.close()
            if pipe:
                monkeypatch.delenv("LISA_PIPELINE")
        np.testing.assert_array_equal(imgs[0], imgs[1])
        np.testing.assert_array_equal(imgs[0], imgs[2])


def test_subframe_additivity(rt, cornell):
    """render(0..3) == equal-weight merge of render(0..1) and render(2..3): the identity the multi-GPU
    partition relies on (lisa_b200/dist.py)."""
    sc = resized(cornell, 64)
    R = rt.Renderer.from_scene(sc)
    R.render_subframes(0, 4, 2)
    full = R.read_accum()
    R.reset()
    R.render_subframes(0, 2, 2)
    a = R.read_accum()
    R.reset()
    R.render_subframes(2, 2, 2)
    b = R.read_accum()
    np.testing.assert_allclose(full[..., :3], 0.5 * (a[..., :3] + b[..., :3]), rtol=2e-6, atol=1e-7)
    # accumulating across calls is the same as one call
    R.reset()
    R.render_subframes(0, 2, 2)
    R.render_subframes(2, 2, 2)
    np.testing.assert_allclose(R.read_accum(), full, rtol=2e-6, atol=1e-7)
    assert R.stats()["subframes_accumulated"] == 4


def test_full_resolution_c2_smoke_properties(rt, cornell):
    """2000x2000 (BASELINE configs[1] resolution), few spp: energy bounds, black border, ray budget."""
    sc = resized(cornell, 2000)
    R = rt.Renderer.from_scene(sc)
    R.render_subframes(0, 1, 2)
    acc, st = R.read_accum(), R.stats()
    assert np.isfinite(acc).all() and (acc[..., :3] >= 0).all()
    # radiance <= sum over <=7 bounces of emission(1) * brdf(<=1/pi) + one emitter hit  (Q3)
    assert acc[..., :3].max() <= 1.0 + 7 / np.pi + 1e-3
    assert acc[:, :60, :3].max() == 0 and acc[:, -60:, :3].max() == 0  # camera sees past the box edges: miss = 0
    assert st["last_samples"] == 2000 * 2000 * 2
    assert st["last_radiance_rays"] <= 7 * st["last_samples"]
    assert st["last_shadow_rays"] <= 30 * st["last_radiance_rays"]
    np.testing.assert_allclose(acc[..., :3].reshape(-1, 3).mean(0), [0.08569, 0.09205, 0.04044], rtol=0.02)  # OptiX 1024-spp mean (512x512)


def test_zero_bounces_and_one_bounce(rt, cornell):
    R = rt.Renderer.from_scene(resized(cornell, 32, bounces=0))
    R.render_subframes(0, 1, 2)
    assert float(R.read_accum()[..., :3].max()) == 0.0
    R = rt.Renderer.from_scene(resized(cornell, 64, bounces=1))
    R.render_subframes(0, 1, 4)
    a = R.read_accum()[..., :3]
    assert a.max() <= 1.0 + 1 / np.pi + 1e-4 and a[53, 32].min() > 0.99  # the light itself


def test_rgba8_and_ppm(rt, orc, cornell, tmp_path):
    import ctypes
    R = rt.Renderer.from_scene(resized(cornell, 80, 48))
    R.render_subframes(0, 1, 8)
    acc, px = R.read_accum(), R.read_rgba8()
    exp = np.zeros_like(px)
    L = orc.lib()
    for y in range(48):
        for x in range(80):
            o = (ctypes.c_uint8 * 4)()
            L.orc_make_color((ctypes.c_float * 3)(*acc[y, x, :3]), o)
            exp[y, x] = list(o)
    diff = np.abs(px.astype(int) - exp.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01  # fast-math powf may move a value across one boundary
    assert (px[..., 3] == 255).all()
    p = str(tmp_path / "o.ppm")
    R.write_ppm(p)
    raw = open(p, "rb").read()
    hdr = b"P6\n80 48\n255\n"
    assert raw.startswith(hdr) and len(raw) == len(hdr) + 80 * 48 * 3
    img = np.frombuffer(raw[len(hdr):], np.uint8).reshape(48, 80, 3)
    np.testing.assert_array_equal(img, px[::-1, :, :3])  # rows reversed, alpha dropped


def test_write_image_follows_the_extension(rt, cornell, tmp_path):
    """save_image -> sutil::saveImage (render.cc:9-17, sutil.cpp:523-688): "ppm" writes the flipped P6, "png" an 8-bit RGBA
    PNG of the same frame (flipped, alpha 255); "exr", unknown extensions and names shorter than 5 characters fail with
    the reference's messages instead of writing a PPM under another name."""
    from PIL import Image
    R = rt.Renderer.from_scene(resized(cornell, 72, 40))
    R.render_subframes(0, 1, 4)
    px = R.read_rgba8()
    p = str(tmp_path / "o.ppm")
    R.write_image(p)
    q = str(tmp_path / "ref.ppm")
    R.write_ppm(q)
    assert open(p, "rb").read() == open(q, "rb").read()
    for name in ("o.png", "o.PNG"):
        p = str(tmp_path / name)
        R.write_image(p)
        im = Image.open(p)
        assert im.mode == "RGBA" and im.size == (72, 40)
        np.testing.assert_array_equal(np.asarray(im), px[::-1])
    for name, msg in (("o.exr", "saving of uchar4 images to EXR not implemented yet"), ("o.jpg", "Failed unsupported filetype 'jpg'"),
                      ("a.pp", "Failed unsupported filetype '.pp'"), ("ppm", "Failed to determine filename extension")):
        with pytest.raises(rt.LisaError, match=msg):
            R.write_image(str(tmp_path / name) if len(name) > 3 else name)
    R.close()


def test_host_render_and_display(rt, frontend, tmp_path, capfd):
    """render()/display() of the reference (render.cc:133-148, 75-131) through liblisa_host.so."""
    os.makedirs(os.path.join(ROOT, "out"), exist_ok=True)
    sc = frontend.Scene("scenes/cornell_tiny.rto")
    d = sc.as_dict()
    out = os.path.join(ROOT, d["output_image"])
    imgs = []
    for progressive in (False, True):
        if os.path.exists(out):
            os.remove(out)
        R = rt.Renderer.from_scene(d)
        frontend.host_render(R, sc, progressive)
        so = capfd.readouterr().out
        assert "Rendering finished in " in so
        assert os.path.exists(out)
        imgs.append(R.read_accum())
        assert R.stats()["samples"] == 64 * 64 * 16
    # -s (1 x 16 spp, subframe 0) and -d (1 x 16 spp, subframe 0) coincide for num_samples = 16
    np.testing.assert_array_equal(imgs[0], imgs[1])


def test_accumulator_checkpoint(rt, cornell, tmp_path):
    """lisa_save_accum / lisa_load_accum: a render resumed in a NEW context from a checkpoint is bit-identical to the
    uninterrupted one; a checkpoint of another image size is refused."""
    sc = resized(cornell, 48, 40)
    R = rt.Renderer.from_scene(sc)
    for f in range(4):
        R.render_subframes(f, 1, 3)
    full = R.read_accum()
    R.close()
    ck = str(tmp_path / "acc.bin")
    R = rt.Renderer.from_scene(sc)
    R.render_subframes(0, 2, 3)
    R.save_accum(ck)
    R.close()
    R = rt.Renderer.from_scene(sc)
    assert R.load_accum(ck) == 2 and R.stats()["subframes_accumulated"] == 2
    R.render_subframes(2, 1, 3)
    R.render_subframes(3, 1, 3)
    np.testing.assert_array_equal(R.read_accum(), full)
    assert R.stats()["subframes_accumulated"] == 4
    R.close()
    R = rt.Renderer.from_scene(resized(cornell, 40, 48))
    with pytest.raises(rt.LisaError, match="holds a 48x40 image"):
        R.load_accum(ck)
    R.close()


def test_pfm_output(rt, cornell, tmp_path):
    R = rt.Renderer.from_scene(resized(cornell, 40, 24))
    R.render_subframes(0, 1, 4)
    p = str(tmp_path / "o.pfm")
    R.write_pfm(p)
    raw = open(p, "rb").read()
    hdr = b"PF\n40 24\n-1.0\n"
    assert raw.startswith(hdr)
    img = np.frombuffer(raw[len(hdr):], "<f4").reshape(24, 40, 3)
    np.testing.assert_array_equal(img, R.read_accum()[..., :3])


def test_ggx_variant_matches_its_oracle(built, orc, tmp_path):
    """Second BSDF (bsdf/ggx.cuh) behind the compile-time seam, checked like the first one: the GGX build of the library
    (liblisa_rt_ggx.so) against the oracle switched to the same BSDF (oracle/cpu_ref.c: ggx_bounce / ggx_brdf), same
    seeds, same estimator — the Lambertian bars of test_matches_oracle_same_seeds (a pixel whose decisions all agree
    reproduces the oracle to 1e-4; ray counts within 0.5 % / 2 %).  The GGX image must also differ from the Lambertian one
    where materials are rough-specular, and leave emitter/miss pixels untouched."""
    import subprocess, sys
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lisa_b200"), "BSDF=ggx", "variant"])
    code = ("import json, numpy as np, lisa_b200.frontend as fe, lisa_b200.rt as rt\n"
            "sc = fe.parse_scene('scenes/cornell_c1.rto'); sc['width'] = sc['height'] = 64\n"
            "for m in sc['materials']:\n"
            "    if 'roughness' in m: m['roughness'] = 0.4\n"
            "sc.pop('materials_packed')\n"
            "res = {}\n"
            "for spp in (1, 8):\n"
            "    R = rt.Renderer.from_scene(sc); R.render_subframes(0, 1, spp); res['a' + str(spp)] = R.read_accum(); st = R.stats()\n"
            "    res['rays' + str(spp)] = np.array([st['last_radiance_rays'], st['last_shadow_rays']])\n"
            "np.savez(r'OUTFILE', **res)\n")
    imgs = []
    for lib in ("liblisa_rt.so", "liblisa_rt_ggx.so"):
        out = str(tmp_path / (lib + ".npz"))
        env = dict(os.environ, LISA_RT_LIB=os.path.join(ROOT, "lisa_b200", lib))
        subprocess.check_call([sys.executable, "-c", code.replace("OUTFILE", out)], cwd=ROOT, env=env)
        imgs.append(np.load(out))
    lam, ggx = imgs
    # the oracle with the same materials
    from oracle import scene_py
    sc = scene_py.parse_scene("scenes/cornell_c1.rto")
    mats = [dict(m, roughness=0.4) if "roughness" in m else m for m in sc["materials"]]
    mp = b"".join(scene_py.pack_material(roughness=m.get("roughness", 0.0), alpha=m["alpha"], n=m.get("n", 0.0), diffuse=m.get("diffuse", (0, 0, 0)),
                                         emit=m["emit"], emission=m.get("emission", (0, 0, 0))) for m in mats)
    S = orc.Scene(sc["vertices"], sc["normals"], sc["mat_indices"], mp)
    cam = sc["camera"]
    try:
        for name, img in (("lambertian", lam), ("ggx", ggx)):
            orc.set_bsdf(name)
            for spp, frac in ((1, 0.995), (8, 0.97)):
                ref, cnt = S.render(cam["eye"], cam["look_at"], cam["fov"], 64, 64, 7, spp)
                acc = img["a%d" % spp]
                assert (np.abs(acc[..., :3] - ref[..., :3]).max(axis=2) < 1e-4).mean() >= frac, (name, spp)
                np.testing.assert_allclose(acc[..., :3].mean(), ref[..., :3].mean(), rtol=0.01)
                rays = img["rays%d" % spp]
                assert abs(int(rays[0]) - cnt["radiance_rays"]) <= 0.005 * cnt["radiance_rays"] + 8
                assert abs(int(rays[1]) - cnt["shadow_rays"]) <= 0.02 * cnt["shadow_rays"] + 8
    finally:
        orc.set_bsdf("lambertian")
    a, b = lam["a8"], ggx["a8"]
    assert np.isfinite(b).all() and (b[..., :3] >= 0).all() and b[..., :3].max() < 50
    assert np.abs(a[..., :3] - b[..., :3]).mean() > 1e-3          # a different BSDF gives a different image
    np.testing.assert_array_equal(a[:, :2], b[:, :2])             # border pixels miss everything: black in both
    assert abs(a[53, 32, :3].min() - b[53, 32, :3].min()) < 0.3  # the light is seen directly in both
