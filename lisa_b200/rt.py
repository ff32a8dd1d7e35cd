"""ctypes binding of liblisa_rt.so — the C ABI declared in include/lisa_rt.h.

Nothing here computes: every call goes to the CUDA library.  Importing this module without the built
library raises (the product has no CPU or PyTorch fallback).
"""
import ctypes
import os

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LISA_RT_LIB") or os.path.join(_DIR, "liblisa_rt.so")  # LISA_RT_LIB: developer override

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "lisa_b200: %s is missing — build it with `make -C lisa_b200` (or __graft_entry__.build()); "
        "there is no CPU fallback" % LIB_PATH)

_L = ctypes.CDLL(LIB_PATH)

LISA_OK = 0
SHADOW_CLOSEST, SHADOW_FIRST_FOUND = 0, 1
BVH_WIDE8, BVH_BINARY = 0, 1
FLAG_PROFILE_STAGES = 1
FLAG_LBVH = 2
FLAG_NO_CULL = 4
FLAG_WAVEFRONT = 8
FLAG_SPLIT_TRIANGLES = 16


class Material(ctypes.Structure):
    _fields_ = [("roughness", ctypes.c_float), ("alpha", ctypes.c_float), ("n", ctypes.c_float),
                ("diffuse_color", ctypes.c_float * 3), ("emit", ctypes.c_uint8), ("_pad", ctypes.c_uint8 * 3),
                ("emission_color", ctypes.c_float * 3)]


class Camera(ctypes.Structure):
    _fields_ = [("eye", ctypes.c_float * 3), ("look_at", ctypes.c_float * 3), ("fov", ctypes.c_float)]


class SceneDesc(ctypes.Structure):
    _fields_ = [("vertices", ctypes.c_void_p), ("normals", ctypes.c_void_p), ("materials", ctypes.c_void_p),
                ("mat_indices", ctypes.c_void_p), ("num_vertices", ctypes.c_int32), ("num_materials", ctypes.c_int32),
                ("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("camera", Camera),
                ("num_samples", ctypes.c_uint32), ("num_bounces", ctypes.c_uint32), ("output_image", ctypes.c_char_p)]


class Options(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("device", ctypes.c_int32), ("shadow_mode", ctypes.c_uint32),
                ("bvh_kind", ctypes.c_uint32), ("max_chains", ctypes.c_uint32), ("flags", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("num_triangles", ctypes.c_uint32),
                ("num_emitter_triangles", ctypes.c_uint32), ("bvh_nodes", ctypes.c_uint32),
                ("bvh_emitter_nodes", ctypes.c_uint32), ("bvh_bytes", ctypes.c_uint64), ("triangle_bytes", ctypes.c_uint64),
                ("upload_ms", ctypes.c_float), ("bvh_build_ms", ctypes.c_float),
                ("samples", ctypes.c_uint64), ("radiance_rays", ctypes.c_uint64), ("shadow_rays", ctypes.c_uint64),
                ("null_directions", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("iterations", ctypes.c_uint64),
                ("render_ms", ctypes.c_double), ("last_render_ms", ctypes.c_double),
                ("last_samples", ctypes.c_uint64), ("last_radiance_rays", ctypes.c_uint64),
                ("last_shadow_rays", ctypes.c_uint64), ("last_kernel_launches", ctypes.c_uint64),
                ("last_extend_ms", ctypes.c_double), ("last_shadow_ms", ctypes.c_double),
                ("state_bytes", ctypes.c_uint64), ("subframes_accumulated", ctypes.c_uint32), ("num_references", ctypes.c_uint32),
                ("last_extend_launches", ctypes.c_uint64), ("last_shadow_launches", ctypes.c_uint64),
                ("last_shadow_jobs", ctypes.c_uint64), ("nodes_visited", ctypes.c_uint64),
                ("triangles_tested", ctypes.c_uint64), ("last_nodes_visited", ctypes.c_uint64),
                ("last_triangles_tested", ctypes.c_uint64), ("shadow_culled", ctypes.c_uint64),
                ("last_shadow_culled", ctypes.c_uint64), ("build_sort_ms", ctypes.c_float), ("build_hierarchy_ms", ctypes.c_float),
                ("build_collapse_ms", ctypes.c_float), ("build_pack_ms", ctypes.c_float),
                ("bvh_sah_nodes_per_ray", ctypes.c_float), ("pool_flavour", ctypes.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


assert ctypes.sizeof(Material) == 40

_vp = ctypes.c_void_p
_L.lisa_last_error.restype = ctypes.c_char_p
_L.lisa_version.restype = ctypes.c_int
_L.lisa_create.argtypes = [ctypes.POINTER(SceneDesc), ctypes.POINTER(Options), ctypes.POINTER(_vp)]
_L.lisa_destroy.argtypes = [_vp]
_L.lisa_destroy.restype = None
_L.lisa_render_subframes.argtypes = [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
_L.lisa_reset_accum.argtypes = [_vp]
_L.lisa_read_accum.argtypes = [_vp, _vp]
_L.lisa_read_rgba8.argtypes = [_vp, _vp]
_L.lisa_write_ppm.argtypes = [_vp, ctypes.c_char_p]
_L.lisa_write_pfm.argtypes = [_vp, ctypes.c_char_p]
_L.lisa_write_image.argtypes = [_vp, ctypes.c_char_p]
_L.lisa_save_accum.argtypes = [_vp, ctypes.c_char_p]
_L.lisa_load_accum.argtypes = [_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32)]
_L.lisa_get_stats.argtypes = [_vp, ctypes.POINTER(Stats)]
_L.lisa_save_bvh.argtypes = [_vp, ctypes.c_char_p]
_L.lisa_create_from_bvh.argtypes = [ctypes.POINTER(SceneDesc), ctypes.POINTER(Options), ctypes.c_char_p, ctypes.POINTER(_vp)]
_L.lisa_accum_add_peer.argtypes = [_vp, _vp]
_L.lisa_accum_note_merged.argtypes = [_vp, ctypes.c_uint32, ctypes.c_uint64]
_L.lisa_multi_create.argtypes = [ctypes.POINTER(SceneDesc), ctypes.POINTER(Options), ctypes.c_int, ctypes.POINTER(_vp)]
_L.lisa_multi_destroy.argtypes = [_vp]
_L.lisa_multi_destroy.restype = None
_L.lisa_multi_num_gpus.argtypes = [_vp]
_L.lisa_multi_root.argtypes = [_vp]
_L.lisa_multi_root.restype = _vp
_L.lisa_multi_ctx.argtypes = [_vp, ctypes.c_int]
_L.lisa_multi_ctx.restype = _vp
_L.lisa_multi_backend.argtypes = [_vp]
_L.lisa_multi_backend.restype = ctypes.c_char_p
_L.lisa_multi_reset_accum.argtypes = [_vp]
_L.lisa_multi_render_subframes.argtypes = [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
_L.lisa_multi_render_samples.argtypes = [_vp, ctypes.c_uint32, ctypes.c_uint32]
_L.lisa_multi_last_times.argtypes = [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
_L.lisa_accum_device_ptr.argtypes = [_vp]
_L.lisa_accum_device_ptr.restype = _vp
_L.lisa_accum_bytes.argtypes = [_vp]
_L.lisa_accum_bytes.restype = ctypes.c_size_t
_L.lisa_device.argtypes = [_vp]
_L.lisa_sync.argtypes = [_vp]
_L.lisa_trace_closest.argtypes = [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, _vp, _vp]
_L.lisa_trace_shadow.argtypes = [_vp, _vp, _vp, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, _vp, _vp]
_L.lisa_primary_rays.argtypes = [_vp, ctypes.c_uint32, _vp, _vp]
_L.lisa_kat_eval.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint32, _vp, _vp, _vp, _vp]

EXPORTS = ["lisa_create", "lisa_destroy", "lisa_last_error", "lisa_version", "lisa_render_subframes",
           "lisa_reset_accum", "lisa_read_accum", "lisa_read_rgba8", "lisa_write_ppm", "lisa_write_pfm", "lisa_write_image", "lisa_save_accum", "lisa_load_accum", "lisa_get_stats", "lisa_save_bvh", "lisa_create_from_bvh",
           "lisa_accum_add_peer", "lisa_accum_note_merged", "lisa_multi_create", "lisa_multi_destroy", "lisa_multi_num_gpus", "lisa_multi_root",
           "lisa_multi_ctx", "lisa_multi_backend", "lisa_multi_reset_accum", "lisa_multi_render_subframes", "lisa_multi_render_samples",
           "lisa_multi_last_times", "lisa_accum_device_ptr", "lisa_accum_bytes", "lisa_device", "lisa_sync", "lisa_trace_closest",
           "lisa_trace_shadow", "lisa_primary_rays", "lisa_kat_eval", "lisa_debug_sort_pairs", "lisa_debug_scan_compact"]


class LisaError(RuntimeError):
    def __init__(self, code):
        self.code = code
        super().__init__("lisa_rt error %d: %s" % (code, _L.lisa_last_error().decode("utf-8", "replace")))


def _check(rc):
    if rc != LISA_OK:
        raise LisaError(rc)


def library():
    return _L


def pack_materials(mats):
    """list of dicts (emit, alpha, diffuse, roughness | n, emission) -> ctypes array of lisa_material."""
    arr = (Material * max(len(mats), 1))()
    for i, m in enumerate(mats):
        arr[i].roughness = m.get("roughness", 0.0)
        arr[i].alpha = m.get("alpha", 1.0)
        arr[i].n = m.get("n", 0.0)
        arr[i].diffuse_color[:] = m.get("diffuse", (0.0, 0.0, 0.0))
        arr[i].emit = 1 if m.get("emit") else 0
        arr[i].emission_color[:] = m.get("emission", (0.0, 0.0, 0.0))
    return arr


class Renderer:
    """One lisa_ctx: OptixWrapper + render() of the reference behind the C ABI."""

    def __init__(self, vertices, normals, mat_indices, materials, width, height, eye, look_at, fov, num_samples=1,
                 num_bounces=7, output_image=None, device=-1, shadow_mode=SHADOW_CLOSEST, bvh_kind=BVH_WIDE8, max_chains=0,
                 flags=0, _multi_gpus=None, _bvh_file=None):
        self._v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self._n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        self._m = np.ascontiguousarray(mat_indices, dtype=np.int32).reshape(-1)
        if isinstance(materials, (bytes, bytearray)):
            nm = len(materials) // 40
            self._mats = (Material * max(nm, 1)).from_buffer_copy(bytes(materials).ljust(40, b"\0"))
        else:
            nm = len(materials)
            self._mats = pack_materials(materials)
        sd = SceneDesc()
        sd.vertices = self._v.ctypes.data
        sd.normals = self._n.ctypes.data
        sd.materials = ctypes.cast(self._mats, _vp)
        sd.mat_indices = self._m.ctypes.data
        sd.num_vertices = self._v.shape[0]
        sd.num_materials = nm
        sd.width, sd.height = width, height
        sd.camera.eye[:] = [float(x) for x in eye]
        sd.camera.look_at[:] = [float(x) for x in look_at]
        sd.camera.fov = float(fov)
        sd.num_samples, sd.num_bounces = num_samples, num_bounces
        sd.output_image = output_image.encode() if output_image else None
        opt = Options(ctypes.sizeof(Options), device, shadow_mode, bvh_kind, max_chains, flags)
        self.width, self.height = width, height
        self.num_samples, self.num_bounces = num_samples, num_bounces
        self._h = _vp()
        self._m = _vp()
        if _bvh_file is not None:   # a BVH serialised by save_bvh(): the geometry arrays may be empty
            _check(_L.lisa_create_from_bvh(ctypes.byref(sd), ctypes.byref(opt), _bvh_file.encode(), ctypes.byref(self._h)))
        elif _multi_gpus is None:
            _check(_L.lisa_create(ctypes.byref(sd), ctypes.byref(opt), ctypes.byref(self._h)))
        else:   # MultiRenderer: one context per GPU behind a lisa_multi; self._h is the root's (borrowed)
            _check(_L.lisa_multi_create(ctypes.byref(sd), ctypes.byref(opt), _multi_gpus, ctypes.byref(self._m)))
            self._h = _vp(_L.lisa_multi_root(self._m))

    @classmethod
    def from_scene(cls, sc, **kw):
        """sc: dict with vertices, normals, mat_indices, materials(_packed), width, height, camera, num_samples, num_bounces."""
        cam = sc["camera"]
        args = dict(width=sc["width"], height=sc["height"], eye=cam["eye"], look_at=cam["look_at"], fov=cam["fov"],
                    num_samples=sc["num_samples"], num_bounces=sc["num_bounces"], output_image=sc.get("output_image"))
        args.update(kw)
        mats = sc["materials_packed"] if "materials_packed" in sc else sc["materials"]
        return cls(sc["vertices"], sc["normals"], sc["mat_indices"], mats, **args)

    @classmethod
    def from_bvh(cls, sc, path, with_geometry=False, **kw):
        """A context from a BVH file written by save_bvh(): materials, camera and size from `sc`; no soup upload, no build."""
        cam = sc["camera"]
        args = dict(width=sc["width"], height=sc["height"], eye=cam["eye"], look_at=cam["look_at"], fov=cam["fov"],
                    num_samples=sc["num_samples"], num_bounces=sc["num_bounces"], output_image=sc.get("output_image"))
        args.update(kw)
        mats = sc["materials_packed"] if "materials_packed" in sc else sc["materials"]
        z = np.zeros((0, 3), np.float32)
        geo = (sc["vertices"], sc["normals"], sc["mat_indices"]) if with_geometry else (z, z, np.zeros(0, np.int32))
        return cls(*geo, mats, _bvh_file=path, **args)

    def save_bvh(self, path):
        _check(_L.lisa_save_bvh(self._h, path.encode()))

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            _L.lisa_multi_destroy(self._m)
            self._m, self._h = _vp(), _vp()
        elif getattr(self, "_h", None) and self._h.value:
            _L.lisa_destroy(self._h)
            self._h = _vp()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def render_subframes(self, first=0, count=1, spp=None):
        _check(_L.lisa_render_subframes(self._h, first, count, self.num_samples if spp is None else spp))

    def render(self):
        """The reference's `-s` path: one subframe (index 0) of num_samples spp (render.cc:133-148)."""
        self.reset()
        self.render_subframes(0, 1, self.num_samples)

    def reset(self):
        _check(_L.lisa_reset_accum(self._h))

    def read_accum(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        assert out.dtype == np.float32 and out.size == self.height * self.width * 4 and out.flags["C_CONTIGUOUS"]
        _check(_L.lisa_read_accum(self._h, out.ctypes.data))
        return out

    def read_rgba8(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        _check(_L.lisa_read_rgba8(self._h, out.ctypes.data))
        return out

    def write_ppm(self, path=None):
        _check(_L.lisa_write_ppm(self._h, path.encode() if path else None))

    def write_image(self, path=None):
        """save_image() of the reference: the format follows the extension (ppm, png; anything else is an error)."""
        _check(_L.lisa_write_image(self._h, path.encode() if path else None))

    def write_pfm(self, path):
        _check(_L.lisa_write_pfm(self._h, path.encode()))

    def save_accum(self, path):
        _check(_L.lisa_save_accum(self._h, path.encode()))

    def load_accum(self, path):
        """Replaces the accumulators with a checkpoint; returns the number of subframes it holds."""
        n = ctypes.c_uint32(0)
        _check(_L.lisa_load_accum(self._h, path.encode(), ctypes.byref(n)))
        return n.value

    def stats(self):
        s = Stats()
        s.struct_size = ctypes.sizeof(Stats)
        _check(_L.lisa_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def accum_add_peer(self, other):
        """self += other (another Renderer, possibly on another GPU of this process)."""
        _check(_L.lisa_accum_add_peer(self._h, other._h))

    def accum_note_merged(self, subframes, samples):
        """Bookkeeping after the caller reduced other contexts' accumulators into this one's buffer (e.g. NCCL)."""
        _check(_L.lisa_accum_note_merged(self._h, subframes, samples))

    def accum_device_ptr(self):
        return _L.lisa_accum_device_ptr(self._h)

    def accum_bytes(self):
        return _L.lisa_accum_bytes(self._h)

    def device(self):
        return _L.lisa_device(self._h)

    def sync(self):
        _check(_L.lisa_sync(self._h))

    def trace_closest(self, org, dirs, tmin=1e-4, tmax=1e16):
        org = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        n = org.shape[0]
        prim = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        _check(_L.lisa_trace_closest(self._h, org.ctypes.data, dirs.ctypes.data, n, tmin, tmax, prim.ctypes.data, t.ctypes.data))
        return prim, t

    def trace_shadow(self, org, dirs, tmin=1e-4, tmax=1e16):
        org = np.ascontiguousarray(org, dtype=np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        n = org.shape[0]
        oc = np.empty(n, dtype=np.int32)
        light = np.empty(n, dtype=np.int32)
        _check(_L.lisa_trace_shadow(self._h, org.ctypes.data, dirs.ctypes.data, n, tmin, tmax, oc.ctypes.data, light.ctypes.data))
        return oc, light

    def primary_rays(self, subframe=0):
        d = np.empty((self.height, self.width, 3), dtype=np.float32)
        s = np.empty((self.height, self.width), dtype=np.uint32)
        _check(_L.lisa_primary_rays(self._h, subframe, d.ctypes.data, s.ctypes.data))
        return d, s


class MultiRenderer(Renderer):
    """lisa_multi: ONE process, one context per GPU, subframes partitioned over them and combined by one ncclReduce
    (include/lisa_rt.h).  Image read-back and statistics go through the root context (GPU 0) with Renderer's methods."""

    def __init__(self, *a, num_gpus=0, **kw):
        kw.pop("device", None)
        super().__init__(*a, _multi_gpus=num_gpus, **kw)

    @classmethod
    def from_scene(cls, sc, num_gpus=0, **kw):
        cam = sc["camera"]
        args = dict(width=sc["width"], height=sc["height"], eye=cam["eye"], look_at=cam["look_at"], fov=cam["fov"],
                    num_samples=sc["num_samples"], num_bounces=sc["num_bounces"], output_image=sc.get("output_image"))
        args.update(kw)
        mats = sc["materials_packed"] if "materials_packed" in sc else sc["materials"]
        return cls(sc["vertices"], sc["normals"], sc["mat_indices"], mats, num_gpus=num_gpus, **args)

    def num_gpus(self):
        return _L.lisa_multi_num_gpus(self._m)

    def backend(self):
        return _L.lisa_multi_backend(self._m).decode()

    def reset(self):
        _check(_L.lisa_multi_reset_accum(self._m))

    def render_subframes(self, first=0, count=1, spp=None):
        _check(_L.lisa_multi_render_subframes(self._m, first, count, self.num_samples if spp is None else spp))

    def render_samples(self, first=0, num_samples=None):
        """`-s` over G GPUs: num_samples split into G subframes of floor/ceil(N / G) spp, exactly N in total."""
        _check(_L.lisa_multi_render_samples(self._m, first, self.num_samples if num_samples is None else num_samples))

    def last_times(self):
        a, b = ctypes.c_double(0), ctypes.c_double(0)
        _L.lisa_multi_last_times(self._m, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value


_L.lisa_debug_sort_pairs.argtypes = [ctypes.c_int, _vp, _vp, ctypes.c_uint32]
_L.lisa_debug_scan_compact.argtypes = [ctypes.c_int, _vp, _vp, ctypes.c_uint32, _vp, _vp]


def debug_sort_pairs(keys, vals, device=-1):
    k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    _check(_L.lisa_debug_sort_pairs(device, k.ctypes.data, v.ctypes.data, k.shape[0]))
    return k, v


def debug_scan_compact(a, c, device=-1):
    a = np.ascontiguousarray(a, dtype=np.uint32).copy()
    c = np.ascontiguousarray(c, dtype=np.int32).copy()
    tot, kept = ctypes.c_uint32(0), ctypes.c_uint32(0)
    _check(_L.lisa_debug_scan_compact(device, a.ctypes.data, c.ctypes.data, a.shape[0], ctypes.byref(tot), ctypes.byref(kept)))
    return a, tot.value, c[:kept.value]


_KAT_SHAPES = {0: (0, 2, 0, 1), 1: (0, 1, 3, 1), 2: (3, 1, 3, 1), 3: (2, 0, 1, 0), 4: (8, 0, 3, 0), 5: (7, 1, 3, 1),
               6: (6, 0, 1, 0), 7: (3, 0, 0, 1), 8: (21, 0, 3, 0), 9: (0, 1, 3, 1)}


def kat_eval(what, in_f=None, in_u=None, device=-1):
    """Evaluate a device helper on the GPU (lisa_kat_eval).  Returns (out_f, out_u)."""
    fi, ui, fo, uo = _KAT_SHAPES[what]
    if fi:
        in_f = np.ascontiguousarray(in_f, dtype=np.float32).reshape(-1, fi)
        n = in_f.shape[0]
    if ui:
        in_u = np.ascontiguousarray(in_u, dtype=np.uint32).reshape(-1, ui)
        n = in_u.shape[0]
    out_f = np.zeros((n, max(fo, 1)), dtype=np.float32)
    out_u = np.zeros((n, max(uo, 1)), dtype=np.uint32)
    _check(_L.lisa_kat_eval(device, what, n, in_f.ctypes.data if fi else None, in_u.ctypes.data if ui else None,
                            out_f.ctypes.data, out_u.ctypes.data))
    return out_f, out_u
