"""volren_b200 -- B200 (sm_100a) back end for the VolRen volume path tracer.

The product is libvrb200.so (CUDA kernels behind the C ABI of include/vrb200.h) plus the C++ host under
host/ (Renderer API, CLI, `volpy`). This Python package is the thin ctypes face of the C ABI used by the
tests and bench.py; it raises when the CUDA library is missing (no CPU fallback).
"""
from . import _capi, formats, scene  # noqa: F401
from ._capi import Context, Counters, NanoVDBGridData, Params, VrbError, load_library  # noqa: F401

__all__ = ["Context", "Params", "Counters", "VrbError", "NanoVDBGridData", "load_library", "formats", "scene"]
