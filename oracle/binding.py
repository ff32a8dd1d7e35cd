"""ctypes binding of oracle/liboracle.so (cpu_ref.c).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f = ctypes.c_float
c_u32 = ctypes.c_uint32
c_fp = ctypes.POINTER(ctypes.c_float)
c_u32p = ctypes.POINTER(ctypes.c_uint32)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.orc_tea16.restype = c_u32
        L.orc_tea16.argtypes = [c_u32, c_u32]
        L.orc_lcg.restype = c_u32
        L.orc_lcg.argtypes = [c_u32p]
        L.orc_rnd.restype = c_f
        L.orc_rnd.argtypes = [c_u32p]
        L.orc_fresnel.restype = c_f
        L.orc_fresnel.argtypes = [c_f, c_f]
        L.orc_brdf.restype = c_f
        L.orc_brdf.argtypes = [c_fp, c_fp]
        L.orc_hemisphere.argtypes = [c_fp, c_u32p, c_fp]
        L.orc_refract.argtypes = [c_f, c_fp, c_fp, c_f, c_fp]
        L.orc_bounce.argtypes = [c_fp, c_fp, c_u32p, c_f, c_fp]
        L.orc_barycentric_normal.argtypes = [c_fp, c_fp, c_fp, c_fp]
        L.orc_make_color.argtypes = [c_fp, ctypes.POINTER(ctypes.c_uint8)]
        L.orc_uvw.argtypes = [c_fp, c_fp, c_f, c_f, c_fp, c_fp, c_fp]
        L.orc_scene_create.restype = ctypes.c_void_p
        L.orc_scene_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_int]
        L.orc_scene_destroy.argtypes = [ctypes.c_void_p]
        L.orc_closest_hit.restype = ctypes.c_int
        L.orc_closest_hit.argtypes = [ctypes.c_void_p, c_fp, c_fp, c_f, c_f, c_fp]
        L.orc_render.restype = ctypes.c_int
        L.orc_render.argtypes = [ctypes.c_void_p, c_fp, c_fp, c_f, c_u32, c_u32, c_u32, c_u32, c_u32, c_u32,
                                 ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_primary_ray.argtypes = [c_fp, c_fp, c_f, c_u32, c_u32, c_u32, c_u32, c_u32, c_fp, c_u32p]
        L.orc_set_bsdf.argtypes = [ctypes.c_int]
        L.orc_write_ppm.restype = ctypes.c_int
        L.orc_write_ppm.argtypes = [ctypes.c_char_p, ctypes.c_void_p, c_u32, c_u32]
        _LIB = L
    return _LIB


def set_bsdf(name):
    """Which BSDF the oracle's estimator uses: "lambertian" (the reference's, default) or "ggx" (the product's variant)."""
    lib().orc_set_bsdf({"lambertian": 0, "ggx": 1}[name])


def f3(v):
    return (ctypes.c_float * len(v))(*[float(x) for x in v])


class Scene:
    """Triangle soup + materials held by the oracle (orc_scene)."""

    def __init__(self, vertices, normals, mat_indices, materials_packed):
        self.v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self.n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        self.m = np.ascontiguousarray(mat_indices, dtype=np.int32)
        self.ntris = self.m.shape[0]
        assert self.v.shape[0] == 3 * self.ntris and self.n.shape[0] == 3 * self.ntris
        self.mats = bytes(materials_packed)
        assert len(self.mats) % 40 == 0
        self.h = lib().orc_scene_create(self.v.ctypes.data, self.n.ctypes.data, self.m.ctypes.data, self.ntris,
                                        self.mats, len(self.mats) // 40)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def closest_hit(self, o, d, tmin=1e-4, tmax=1e16):
        t = ctypes.c_float(0)
        p = lib().orc_closest_hit(self.h, f3(o), f3(d), tmin, tmax, ctypes.byref(t))
        return p, t.value

    def render(self, eye, look_at, fov, width, height, bounces, spp, first_subframe=0, subframes=1, pixels=None,
               threads=0):
        """Returns (accum[H,W,4] or [len(pixels),4] float32 mean, counters dict)."""
        cnt = np.zeros(4, dtype=np.uint64)
        if pixels is None:
            acc = np.zeros((height, width, 4), dtype=np.float32)
            pp, npx = None, 0
        else:
            pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
            acc = np.zeros((pixels.shape[0], 4), dtype=np.float32)
            pp, npx = pixels.ctypes.data, pixels.shape[0]
        used = lib().orc_render(self.h, f3(eye), f3(look_at), fov, width, height, bounces, first_subframe, subframes,
                                spp, pp, npx, acc.ctypes.data, cnt.ctypes.data, threads)
        return acc, dict(radiance_rays=int(cnt[0]), shadow_rays=int(cnt[1]), samples=int(cnt[2]),
                         null_dirs=int(cnt[3]), threads=used)
