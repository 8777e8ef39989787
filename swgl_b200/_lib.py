"""Loader for libswgl_b200.so -- the CUDA draw-call path behind the swgl C ABI.

There is no CPU fallback: if the shared library is missing this raises, and if no CUDA
device is present ``glInit`` reports the failure through ``swglGetLastError`` and every
hot-path call is a no-op.
"""
from __future__ import annotations

import ctypes as C
import os

from . import gl as G

HERE = os.path.dirname(os.path.abspath(__file__))
# SWGL_B200_LIB: development override (A/B of two builds on one GPU box, tools/ab_lib.py)
LIB_PATH = os.environ.get("SWGL_B200_LIB") or os.path.join(HERE, "libswgl_b200.so")


class swglStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("draws", "triangles_in", "prims_out", "tested", "shaded", "tile_pairs", "bands")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# extension entry points (include/swgl.h glDrawElements + include/swgl_b200.h)
EXTENSION_EXPORTS = {
    "glDrawElements": (None, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "swglGetDepthPtr": (C.POINTER(C.c_float), []),
    "swglFinish": (None, []),
    "swglGetLastError": (C.c_char_p, []),
    "swglGetStats": (None, [C.POINTER(swglStats)]),
    "swglSetDevice": (None, [C.c_int]),
    "swglGetStream": (C.c_void_p, []),
    "swglGetColorDevicePtr": (C.c_uint64, []),
    "swglPrecompileProgram": (C.c_int, []),
    "swglSetDeviceCount": (None, [C.c_int]),
    "swglGetDepthDevicePtr": (C.c_uint64, []),
    "swglFillFramebuffer": (None, [C.c_uint32, C.c_float]),
    "swglBufferRespecify": (None, [C.c_uint32, C.c_uint32, C.c_void_p]),
    "swglSetStripe": (None, [C.c_uint32, C.c_uint32, C.c_uint32]),
    "swglSetPeerColorTarget": (None, [C.c_uint64]),
    "swglIpcExportColor": (C.c_int, [C.c_void_p]),
    "swglIpcOpen": (C.c_uint64, [C.c_void_p]),
    "swglIpcClose": (None, [C.c_uint64]),
    "swglSetOption": (None, [C.c_char_p, C.c_int64]),
    "swglGetOption": (C.c_int64, [C.c_char_p]),
    "swglDebugShaderIR": (C.c_size_t, [C.c_uint32, C.c_char_p, C.c_size_t]),
    "swglGetShaderCompiled": (C.c_int, [C.c_uint32]),
    "swglFrameSubmit": (C.c_uint64, []),
    "swglFrameWait": (C.POINTER(C.c_uint32), [C.c_uint64]),
    "swglReadPixelsRGBA8": (C.c_int, [C.c_void_p]),
    "swglWritePPM": (C.c_int, [C.c_char_p]),
    "swglHashWords": (C.c_uint64, [C.c_void_p, C.c_uint64]),
    "swglBufferSubData": (None, [C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p]),
    "swglGetBufferDevicePtr": (C.c_uint64, [C.c_uint32]),
    "swglBufferDeviceWritten": (None, [C.c_uint32]),
    "swglSetSharedFrameMirror": (C.c_int, [C.c_void_p, C.c_uint64]),
    "swglHostAlloc": (C.c_void_p, [C.c_uint64, C.c_int]),
    "swglHostFree": (None, [C.c_void_p]),
}

_api = None


def load() -> G.GLApi:
    """Load (once) and return the typed API object of the product library."""
    global _api
    if _api is not None:
        return _api
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m swgl_b200.build` "
            "(there is no CPU fallback for the swgl draw path)")
    lib = C.CDLL(LIB_PATH)
    _api = G.GLApi(lib, EXTENSION_EXPORTS)
    return _api
