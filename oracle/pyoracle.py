"""Python bindings for the CPU oracle -- TEST INFRASTRUCTURE, never imported by the product.

Two checkers live here:

* ``Restatement``  -- oracle/libswgl_oracle.so, the plain-C restatement of the hot path
  (oracle/swgl_oracle.c), portable, counts tested/shaded fragments;
* ``Reference``    -- oracle/_ref/libswgl_ref.so, the unmodified reference compiled from
  /root/reference by oracle/Makefile (binary only; present on the GPU box as a prebuilt file).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libswgl_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libswgl_ref.so")
# the same reference with rsqrt()'s `long` pun made 32-bit (oracle/ref_shim.c): the mip-map row's checker
REF_LOD_SO = os.path.join(HERE, "_ref", "libswgl_ref_lod.so")


def build(quiet: bool = True) -> None:
    """Compile the restatement and (when /root/reference exists) the reference binary."""
    subprocess.run(
        ["make", "-C", HERE, "all"],
        check=True,
        stdout=subprocess.DEVNULL if quiet else None,
    )


class _Target(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("vx", C.c_int32),
        ("vy", C.c_int32),
        ("vw", C.c_uint32),
        ("vh", C.c_uint32),
        ("color", C.POINTER(C.c_uint32)),
        ("depth", C.POINTER(C.c_float)),
    ]


class _Shader(C.Structure):
    _fields_ = [
        ("use_matrix", C.c_int32),
        ("matrix_value", C.c_float * 16),
        ("matrix_transpose", C.c_int32),
        ("stride", C.c_uint32),
        ("pos_offset", C.c_uint32),
        ("pos_size", C.c_int32),
        ("var_offset", C.c_uint32),
        ("var_size", C.c_int32),
        ("var_comps", C.c_int32),
        ("fs_mode", C.c_int32),
        ("tex", C.POINTER(C.c_float)),
        ("tex_w", C.c_int32),
        ("tex_h", C.c_int32),
        ("tex_fpp", C.c_int32),
        ("wrap_s_repeat", C.c_int32),
        ("wrap_t_repeat", C.c_int32),
        ("lod", C.c_int32),
        ("n_mips", C.c_int32),
        ("mip_data", C.POINTER(C.c_float)),
        ("mip_off", C.POINTER(C.c_int32)),
        ("mip_w", C.POINTER(C.c_int32)),
        ("mip_h", C.POINTER(C.c_int32)),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("triangles_in", C.c_uint64),
        ("prims_out", C.c_uint64),
        ("tested", C.c_uint64),
        ("shaded", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def _scene_shader(scene, keep: list) -> _Shader:
    s = _Shader()
    s.use_matrix = 1 if scene.matrix is not None else 0
    if scene.matrix is not None:
        s.matrix_value = (C.c_float * 16)(*[float(x) for x in scene.matrix])
    s.matrix_transpose = 0  # the scenes call glUniformMatrix4fv(loc, 1, GL_FALSE, value)
    s.stride = scene.stride
    (l0, n0, o0), (l1, n1, o1) = scene.attribs
    assert l0 == 0 and l1 == 1
    s.pos_offset, s.pos_size = o0, n0
    s.var_offset, s.var_size = o1, n1
    s.var_comps = n1
    s.fs_mode = 1 if scene.texture is not None else 0
    if scene.texture is not None:
        t8 = np.ascontiguousarray(scene.texture, np.uint8)
        tf = np.empty(t8.size, np.float32)
        keep.append(tf)
        keep.append(t8)
        s.tex = tf.ctypes.data_as(C.POINTER(C.c_float))
        s.tex_h, s.tex_w, s.tex_fpp = t8.shape[0], t8.shape[1], t8.shape[2]
        rep = 1 if scene.tex_wrap == "repeat" else 0
        s.wrap_s_repeat = s.wrap_t_repeat = rep
    return s


class Restatement:
    """The C restatement (port).  Renders a ``swgl_b200.scenes.Scene`` into numpy arrays."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.swglo_clear.argtypes = [C.POINTER(_Target), C.c_uint32] + [C.c_float] * 4
        L.swglo_draw_arrays.argtypes = [
            C.POINTER(_Target), C.POINTER(_Shader), C.c_void_p, C.c_size_t,
            C.c_int32, C.c_uint32, C.POINTER(Stats)]
        L.swglo_draw_elements.argtypes = [
            C.POINTER(_Target), C.POINTER(_Shader), C.c_void_p, C.c_size_t,
            C.c_void_p, C.c_uint32, C.POINTER(Stats)]
        L.swglo_draw_points.argtypes = [C.POINTER(_Target), C.POINTER(_Shader), C.c_void_p, C.c_size_t, C.c_int32, C.c_uint32]
        L.swglo_texels_from_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.swglo_build_mipmaps.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(C.c_size_t)]
        L.swglo_build_mipmaps.restype = C.c_int
        L.swglo_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.swglo_fnv1a64.restype = C.c_uint64
        L.swglo_timed_frame.argtypes = [
            C.POINTER(_Target), C.POINTER(_Shader), C.c_void_p, C.c_size_t, C.c_void_p,
            C.c_int32, C.c_uint32] + [C.c_float] * 4 + [C.c_int, C.POINTER(Stats)]
        L.swglo_timed_frame.restype = C.c_double

    def fnv(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a).view(np.uint32).reshape(-1)
        return int(self.lib.swglo_fnv1a64(a.ctypes.data_as(C.c_void_p), a.size))

    def _prep(self, scene, color, depth, keep, mipmaps=False):
        t = _Target()
        t.width, t.height = scene.width, scene.height
        vp = scene.viewport or (0, 0, scene.width, scene.height)
        t.vx, t.vy, t.vw, t.vh = vp
        t.color = color.ctypes.data_as(C.POINTER(C.c_uint32))
        t.depth = depth.ctypes.data_as(C.POINTER(C.c_float))
        s = _scene_shader(scene, keep)
        if scene.texture is not None:
            tf, t8 = keep[0], keep[1]
            self.lib.swglo_texels_from_u8(t8.ctypes.data_as(C.c_void_p), tf.ctypes.data_as(C.c_void_p), tf.size)
            if mipmaps:
                # glGenerateMipmap + the defined level of detail (swgl_oracle.h): size query, then the levels
                total = C.c_size_t(0)
                n = self.lib.swglo_build_mipmaps(tf.ctypes.data_as(C.c_void_p), s.tex_w, s.tex_h, s.tex_fpp, None, None, None, None, C.byref(total))
                data = np.zeros(max(int(total.value), 1), np.float32)
                off, lw, lh = (np.zeros(max(n, 1), np.int32) for _ in range(3))
                self.lib.swglo_build_mipmaps(tf.ctypes.data_as(C.c_void_p), s.tex_w, s.tex_h, s.tex_fpp, data.ctypes.data_as(C.c_void_p),
                                             off.ctypes.data_as(C.c_void_p), lw.ctypes.data_as(C.c_void_p), lh.ctypes.data_as(C.c_void_p), C.byref(total))
                keep += [data, off, lw, lh]
                s.lod, s.n_mips = 1, n
                s.mip_data = data.ctypes.data_as(C.POINTER(C.c_float))
                s.mip_off = off.ctypes.data_as(C.POINTER(C.c_int32))
                s.mip_w = lw.ctypes.data_as(C.POINTER(C.c_int32))
                s.mip_h = lh.ctypes.data_as(C.POINTER(C.c_int32))
        return t, s

    def render(self, scene, *, clear: bool = True, color=None, depth=None,
               first: int = 0, count: Optional[int] = None, fill=(0, 0.0), mipmaps: bool = False, points=None):
        """Returns (color uint32 [H,W], depth float32 [H,W], stats dict).
        points = (first, count): a glDrawArrays(GL_POINTS, first, count) over the same vertex stream after the triangles.
        mipmaps: glGenerateMipmap on the scene's texture and the DEFINED level of detail (the checker of that mode is
        Reference(defined_rsqrt=True), rendered with mipmaps=True too)."""
        H, W = scene.height, scene.width
        if color is None:
            color = np.full((H, W), fill[0], np.uint32)
        if depth is None:
            depth = np.full((H, W), fill[1], np.float32)
        keep: list = []
        t, s = self._prep(scene, color, depth, keep, mipmaps)
        if clear:
            self.lib.swglo_clear(C.byref(t), 3, *[C.c_float(c) for c in scene.clear_color])
        st = Stats()
        v = np.ascontiguousarray(scene.vertices, np.float32)
        if scene.indices is not None:
            idx = np.ascontiguousarray(scene.indices, np.uint32)
            n = len(idx) if count is None else count
            self.lib.swglo_draw_elements(C.byref(t), C.byref(s), v.ctypes.data_as(C.c_void_p), v.nbytes,
                                         idx[first:].ctypes.data_as(C.c_void_p), n, C.byref(st))
        else:
            n = len(v) if count is None else count
            self.lib.swglo_draw_arrays(C.byref(t), C.byref(s), v.ctypes.data_as(C.c_void_p), v.nbytes,
                                       first, n, C.byref(st))
        if points is not None:
            self.lib.swglo_draw_points(C.byref(t), C.byref(s), v.ctypes.data_as(C.c_void_p), v.nbytes, points[0], points[1])
        return color, depth, st.as_dict()

    def timed_frame(self, scene, count: Optional[int] = None, reps: int = 1):
        """Seconds for clear + draw of the first ``count`` stream vertices (single thread)."""
        H, W = scene.height, scene.width
        color = np.zeros((H, W), np.uint32)
        depth = np.zeros((H, W), np.float32)
        keep: list = []
        t, s = self._prep(scene, color, depth, keep)
        v = np.ascontiguousarray(scene.vertices, np.float32)
        idx = None
        if scene.indices is not None:
            idx = np.ascontiguousarray(scene.indices, np.uint32)
            n = len(idx) if count is None else count
        else:
            n = len(v) if count is None else count
        st = Stats()
        dt = self.lib.swglo_timed_frame(
            C.byref(t), C.byref(s), v.ctypes.data_as(C.c_void_p), v.nbytes,
            idx.ctypes.data_as(C.c_void_p) if idx is not None else None,
            0, n, *[C.c_float(c) for c in scene.clear_color], reps, C.byref(st))
        return float(dt), st.as_dict()


class Reference:
    """The unmodified reference, driven through its own API (swgl.h)."""

    def __init__(self, defined_rsqrt: bool = False):
        so = REF_LOD_SO if defined_rsqrt else REF_SO
        if not os.path.exists(so):
            build()
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        from swgl_b200 import gl as G

        self.G = G
        self.lib = C.CDLL(so)
        self.api = G.GLApi(self.lib)
        self.lib.swglref_depth_ptr.restype = C.POINTER(C.c_float)
        self.lib.swglref_fill.argtypes = [C.c_uint32, C.c_float]
        self.lib.swglref_timed_frame.argtypes = [C.c_uint32, C.c_int32, C.c_uint32, C.POINTER(C.c_double)]
        self.lib.swglref_timed_frame.restype = C.c_double

    def render(self, scene, *, clear: bool = True, first: int = 0, count: Optional[int] = None,
               fill=(0, 0.0), mipmaps: bool = False, points=None):
        """Returns (color uint32 [H,W], depth float32 [H,W]).  Indexed scenes are de-indexed
        on the host first: the reference has no glDrawElements.
        points = (first, count): glDrawArrays(GL_POINTS, first, count) after the triangles."""
        G = self.G
        st = G.setup_scene(self.api, scene, indexed=False)
        if mipmaps:
            self.api.glGenerateMipmap(G.GL_TEXTURE_2D)
        self.lib.swglref_fill(fill[0], fill[1])
        if clear:
            self.api.glClear(3)
        n = st["n_draw"] if count is None else count
        self.api.glDrawArrays(G.GL_TRIANGLES, first, n)
        if points is not None:
            self.api.glDrawArrays(G.GL_POINTS, points[0], points[1])
        H, W = scene.height, scene.width
        col = G.frame_color(self.api, W, H)
        dep = np.ctypeslib.as_array(self.lib.swglref_depth_ptr(), shape=(H, W)).copy()
        return col, dep

    def timed_frame(self, scene, count: Optional[int] = None):
        """Seconds for glClear + glDrawArrays of the first ``count`` vertices (single thread)."""
        G = self.G
        st = G.setup_scene(self.api, scene, indexed=False)
        self.lib.swglref_fill(0, 0.0)
        n = st["n_draw"] if count is None else min(count, st["n_draw"])
        clear_s = C.c_double(0)
        dt = self.lib.swglref_timed_frame(3, 0, n, C.byref(clear_s))
        return float(dt), float(clear_s.value)


def compare(col_a, dep_a, col_b, dep_b) -> dict:
    """Parity counters: coverage/depth bit mismatches, colour word mismatches, max channel delta."""
    da = np.ascontiguousarray(dep_a, np.float32).view(np.uint32)
    db = np.ascontiguousarray(dep_b, np.float32).view(np.uint32)
    nan_a, nan_b = np.isnan(dep_a), np.isnan(dep_b)
    depth_mis = (da != db) & ~(nan_a & nan_b)  # NaN == NaN whatever the payload (A.7)
    cov_mis = (da != 0) != (db != 0)
    ca, cb = np.asarray(col_a, np.uint32), np.asarray(col_b, np.uint32)
    col_mis = ca != cb
    maxd = 0
    if col_mis.any():
        for sh in (24, 16, 8, 0):
            xa = ((ca >> sh) & 0xFF).astype(np.int32)
            xb = ((cb >> sh) & 0xFF).astype(np.int32)
            maxd = max(maxd, int(np.abs(xa - xb).max()))
    return {
        "depth_mismatch": int(depth_mis.sum()),
        "coverage_mismatch": int(cov_mis.sum()),
        "color_mismatch": int(col_mis.sum()),
        "max_channel_delta": maxd,
        "pixels": int(ca.size),
    }
