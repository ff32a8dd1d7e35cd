"""lisa_b200 — B200-native drop-in for the render path of gaetanserre/LiSA.

The product is native: ``liblisa_rt.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/lisa_rt.h``), ``liblisa_host.so`` (C++ scene parser / OBJ loader / render drivers, mirroring
the reference's ``SceneParser``, ``parse_obj``, ``render`` and ``display``) and the ``lisa`` CLI.
This Python package is only the ctypes binding the tests and ``bench.py`` use; it raises at import of
``lisa_b200.rt`` if the CUDA library has not been built (there is no CPU fallback).
"""
__all__ = ["rt", "frontend", "dist"]
