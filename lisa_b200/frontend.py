"""ctypes binding of liblisa_host.so (include/lisa_host.h): the C++ scene front-end and render drivers.

`parse_scene` is SceneParser(path).get_params() of the reference (src/LiSA/include/scene_parser.hh:10-13)
as implemented by lisa_b200/host/scene_parser.cc; nothing is parsed in Python.
"""
import ctypes
import os

import numpy as np

from . import rt

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "liblisa_host.so")
if not os.path.exists(LIB_PATH):
    raise ImportError("lisa_b200: %s is missing — build it with `make -C lisa_b200`" % LIB_PATH)
_H = ctypes.CDLL(LIB_PATH)
_vp = ctypes.c_void_p
_H.lisa_scene_parse.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_vp)]
_H.lisa_scene_free.argtypes = [_vp]
_H.lisa_scene_free.restype = None
_H.lisa_scene_get_desc.argtypes = [_vp]
_H.lisa_scene_get_desc.restype = ctypes.POINTER(rt.SceneDesc)
_H.lisa_scene_num_meshes.argtypes = [_vp]
_H.lisa_scene_mesh_file.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
_H.lisa_scene_mesh_file.restype = ctypes.c_char_p
_H.lisa_scene_material_index.argtypes = [_vp, ctypes.c_char_p]
_H.lisa_host_render.argtypes = [_vp, _vp, ctypes.c_int]
_H.lisa_host_set_obj_cache.argtypes = [ctypes.c_int]
_H.lisa_host_set_obj_cache.restype = None
_H.lisa_host_last_error.restype = ctypes.c_char_p
_H.lisa_host_last_exit_code.restype = ctypes.c_int

EXPORTS = ["lisa_scene_parse", "lisa_scene_free", "lisa_scene_get_desc", "lisa_scene_num_meshes", "lisa_scene_mesh_file",
           "lisa_scene_material_index", "lisa_host_render", "lisa_host_set_obj_cache", "lisa_host_last_error",
           "lisa_host_last_exit_code"]


def set_obj_cache(enabled):
    """Binary soup cache of the OBJ loader (<file>.lisasoup next to each OBJ): True/False, or None = as LISA_OBJ_CACHE says."""
    _H.lisa_host_set_obj_cache(-1 if enabled is None else int(bool(enabled)))


class SceneError(Exception):
    """What the reference prints on stderr before exit(code)."""

    def __init__(self, message, code):
        super().__init__(message)
        self.code = code


class Scene:
    """Owns a parsed scene (SceneParser object on the C++ side)."""

    def __init__(self, path, load_meshes=True):
        self._h = _vp()
        rc = _H.lisa_scene_parse(os.fsencode(path), 1 if load_meshes else 0, ctypes.byref(self._h))
        if rc != 0:
            raise SceneError(_H.lisa_host_last_error().decode("utf-8", "replace"), _H.lisa_host_last_exit_code())
        self.desc = _H.lisa_scene_get_desc(self._h).contents

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _H.lisa_scene_free(self._h)
            self._h = _vp()

    __del__ = close

    def handle(self):
        return self._h

    def mesh_files(self):
        out = []
        for i in range(_H.lisa_scene_num_meshes(self._h)):
            mi = ctypes.c_int(0)
            out.append((_H.lisa_scene_mesh_file(self._h, i, ctypes.byref(mi)).decode(), mi.value))
        return out

    def material_index(self, name):
        return _H.lisa_scene_material_index(self._h, name.encode())

    def as_dict(self):
        d = self.desc
        nv, nm = d.num_vertices, d.num_materials
        f32 = lambda p, n: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n,)).copy() if n else np.zeros(0, np.float32)
        verts = f32(d.vertices, nv * 3).reshape(-1, 3)
        norms = f32(d.normals, nv * 3).reshape(-1, 3)
        nt = nv // 3
        midx = np.ctypeslib.as_array(ctypes.cast(d.mat_indices, ctypes.POINTER(ctypes.c_int32)), shape=(nt,)).copy() if nt else np.zeros(0, np.int32)
        packed = ctypes.string_at(d.materials, 40 * nm) if nm else b""
        mats = []
        marr = ctypes.cast(d.materials, ctypes.POINTER(rt.Material))
        for i in range(nm):
            m = marr[i]
            if m.emit:
                mats.append(dict(emit=True, alpha=m.alpha, emission=tuple(m.emission_color)))
            elif m.alpha < 1.0:
                mats.append(dict(emit=False, alpha=m.alpha, diffuse=tuple(m.diffuse_color), n=m.n))
            else:
                mats.append(dict(emit=False, alpha=m.alpha, diffuse=tuple(m.diffuse_color), roughness=m.roughness))
        return dict(width=d.width, height=d.height, num_samples=d.num_samples, num_bounces=d.num_bounces,
                    output_image=d.output_image.decode("utf-8", "replace") if d.output_image else None,
                    camera=dict(eye=tuple(d.camera.eye), look_at=tuple(d.camera.look_at), fov=d.camera.fov),
                    vertices=verts, normals=norms, mat_indices=midx, materials=mats, materials_packed=packed,
                    mesh_files=self.mesh_files())


def parse_scene(path, load_meshes=True):
    s = Scene(path, load_meshes)
    try:
        return s.as_dict()
    finally:
        s.close()


def host_render(renderer, scene, progressive=False):
    """render() / display() of the reference on an existing Renderer (lisa_ctx) and Scene."""
    rc = _H.lisa_host_render(renderer._h, scene.handle(), 1 if progressive else 0)
    if rc != 0:
        raise RuntimeError(_H.lisa_host_last_error().decode("utf-8", "replace"))
