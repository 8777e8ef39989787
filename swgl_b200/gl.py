"""ctypes mirror of the swgl.h entry points (reference: swgl.h:87-158).

The same thin binding drives three different shared libraries that all export the swgl API:

* ``libswgl_b200.so``  -- the product: CUDA draw-call path behind the C ABI (include/swgl.h);
* ``oracle/_ref/libswgl_ref.so`` -- the unmodified reference, test infrastructure only;
* nothing else: there is no CPU fallback in the product.

Names, argument order and enum values are the reference's own (positional C enum,
swgl.h:40-81 -- *not* Khronos values), so a test written against one library runs
unchanged against the other.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

# --- positional enum of swgl.h:40-81 ---------------------------------------------------------
(
    GL_VERTEX_SHADER,
    GL_FRAGMENT_SHADER,
    GL_COMPILE_STATUS,
    GL_LINK_STATUS,
    GL_ARRAY_BUFFER,
    GL_STATIC_DRAW,
    GL_STREAM_DRAW,
    GL_DYNAMIC_DRAW,
    GL_FLOAT,
    GL_INT,
    GL_UNSIGNED_BYTE,
    GL_DEPTH_COMPONENT,
    GL_DEPTH_STENCIL,
    GL_RED,
    GL_RG,
    GL_RGB,
    GL_RGBA,
    GL_TRIANGLES,
    GL_POINTS,
    GL_LINES,
    GL_REPEAT,
    GL_CLAMP,
    GL_TEXTURE_2D,
    GL_TEXTURE_WRAP_S,
    GL_TEXTURE_WRAP_T,
    GL_TEXTURE0,
    GL_TEXTURE1,
    GL_TEXTURE2,
    GL_TEXTURE3,
    GL_TEXTURE4,
    GL_TEXTURE5,
    GL_TEXTURE6,
    GL_TEXTURE7,
    # extension enumerators appended by include/swgl.h (values after GL_TEXTURE7)
    GL_ELEMENT_ARRAY_BUFFER,
    GL_UNSIGNED_INT,
) = range(35)

GL_COLOR_BUFFER_BIT = 1  # swgl.h:19
GL_DEPTH_BUFFER_BIT = 2  # swgl.h:20
GL_TRUE, GL_FALSE = 1, 0

_u32, _i32, _f32, _u8 = C.c_uint32, C.c_int32, C.c_float, C.c_uint8
_vp, _cp = C.c_void_p, C.c_char_p

# name -> (restype, argtypes): the 36 reference exports
REFERENCE_EXPORTS = {
    "glInit": (None, [_u32, _u32]),
    "glGetFramePtr": (C.POINTER(_u32), []),
    "glCreateShader": (_u32, [_u32]),
    "glShaderSource": (None, [_u32, _cp]),
    "glCompileShader": (None, [_u32]),
    "glDeleteShader": (None, [_u32]),
    "glCreateProgram": (_u32, []),
    "glAttachShader": (None, [_u32, _u32]),
    "glLinkProgram": (None, [_u32]),
    "glUseProgram": (None, [_u32]),
    "glGenVertexArrays": (_u32, [_u32, C.POINTER(_u32)]),
    "glBindVertexArray": (None, [_u32]),
    "glVertexAttribPointer": (None, [_u32, _i32, _u32, _u8, _u32, _vp]),
    "glEnableVertexAttribArray": (None, [_u32]),
    "glGenBuffers": (_u32, [_u32, C.POINTER(_u32)]),
    "glBindBuffer": (None, [_u32, _u32]),
    "glBufferData": (None, [_u32, _u32, _vp, _u32]),
    "glClearColor": (None, [_f32, _f32, _f32, _f32]),
    "glClear": (None, [_u32]),
    "glViewport": (None, [_i32, _i32, _u32, _u32]),
    "glDrawArrays": (None, [_u32, _i32, _u32]),
    "glGenTextures": (None, [_u32, C.POINTER(_u32)]),
    "glActiveTexture": (None, [_u32]),
    "glBindTexture": (None, [_u32, _u32]),
    "glTexParameteri": (None, [_u32, _u32, _u32]),
    "glTexImage2D": (None, [_u32, _i32, _i32, _u32, _u32, _i32, _u32, _u32, _vp]),
    "glGenerateMipmap": (None, [_u32]),
    "glGetUniformLocation": (_i32, [_u32, _cp]),
    "glUniform1f": (None, [_i32, _f32]),
    "glUniform2f": (None, [_i32, _f32, _f32]),
    "glUniform3f": (None, [_i32, _f32, _f32, _f32]),
    "glUniform4f": (None, [_i32, _f32, _f32, _f32, _f32]),
    "glUniform1i": (None, [_i32, _i32]),
    "glUniformMatrix2fv": (None, [_i32, _u32, _u8, C.POINTER(_f32)]),
    "glUniformMatrix3fv": (None, [_i32, _u32, _u8, C.POINTER(_f32)]),
    "glUniformMatrix4fv": (None, [_i32, _u32, _u8, C.POINTER(_f32)]),
}


class GLApi:
    """Typed attribute access to one loaded library exporting the swgl API."""

    def __init__(self, lib: C.CDLL, extra: Optional[dict] = None):
        self.lib = lib
        table = dict(REFERENCE_EXPORTS)
        if extra:
            table.update(extra)
        for name, (res, args) in table.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def add_geometry(gl: GLApi, scene, *, indexed: bool, named: bool = False):
    """A vertex array with its own vertex (and element) buffer holding ``scene``'s geometry; it stays
    bound.  Returns (vao, vbo, ebo or 0, number of vertices / indices to draw).

    With a vertex array bound, glBindBuffer makes the array's OWN Buffer struct the bound buffer and
    overwrites it with a snapshot of the named one (swgl.c:3116-3122), so data specified through it
    never reaches the named buffer and binding the name again wipes it.  ``named=True`` specifies the
    data with no vertex array bound (the named buffers own it); such a set can be re-bound by name
    later, which is what a client that alternates between geometry sets has to do."""
    verts = scene.vertices if indexed else scene.deindexed()
    verts = np.ascontiguousarray(verts, np.float32)
    idx = np.ascontiguousarray(scene.indices, np.uint32) if (indexed and scene.indices is not None) else None
    vao, vbo, ebo = _u32(0), _u32(0), _u32(0)
    if named:
        gl.glBindVertexArray(0)
        gl.glGenBuffers(1, C.byref(vbo))
        gl.glBindBuffer(GL_ARRAY_BUFFER, vbo.value)
        gl.glBufferData(GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), GL_STATIC_DRAW)
        if idx is not None:
            gl.glGenBuffers(1, C.byref(ebo))
            gl.glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, ebo.value)
            gl.glBufferData(GL_ELEMENT_ARRAY_BUFFER, idx.nbytes, _ptr(idx), GL_STATIC_DRAW)
    gl.glGenVertexArrays(1, C.byref(vao))
    gl.glBindVertexArray(vao.value)
    if not named:
        gl.glGenBuffers(1, C.byref(vbo))
    gl.glBindBuffer(GL_ARRAY_BUFFER, vbo.value)
    if not named:
        gl.glBufferData(GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), GL_STATIC_DRAW)
    for loc, n, off in scene.attribs:
        gl.glVertexAttribPointer(loc, n, GL_FLOAT, GL_FALSE, scene.stride, _vp(off))
        gl.glEnableVertexAttribArray(loc)
    n_draw = len(verts)
    if idx is not None:
        if not named:
            gl.glGenBuffers(1, C.byref(ebo))
        gl.glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, ebo.value)
        if not named:
            gl.glBufferData(GL_ELEMENT_ARRAY_BUFFER, idx.nbytes, _ptr(idx), GL_STATIC_DRAW)
        n_draw = len(idx)
    return vao.value, vbo.value, ebo.value, n_draw


def setup_scene(gl: GLApi, scene, *, indexed: bool, init: bool = True):
    """Issue every cold-path call for ``scene``; returns a dict with the program id etc.

    ``indexed=False`` uploads the de-indexed vertex stream (what the reference can draw:
    it has no glDrawElements).
    """
    if init:
        gl.glInit(scene.width, scene.height)
    vs = gl.glCreateShader(GL_VERTEX_SHADER)
    gl.glShaderSource(vs, scene.vs.encode())
    gl.glCompileShader(vs)
    fs = gl.glCreateShader(GL_FRAGMENT_SHADER)
    gl.glShaderSource(fs, scene.fs.encode())
    gl.glCompileShader(fs)
    prog = gl.glCreateProgram()
    gl.glAttachShader(prog, vs)
    gl.glAttachShader(prog, fs)
    gl.glLinkProgram(prog)
    gl.glUseProgram(prog)

    vao, vbo, ebo, n_draw = add_geometry(gl, scene, indexed=indexed)

    if scene.texture is not None:
        tex = _u32(0)
        gl.glGenTextures(1, C.byref(tex))
        gl.glActiveTexture(GL_TEXTURE0)
        gl.glBindTexture(GL_TEXTURE_2D, tex.value)
        wrap = GL_REPEAT if scene.tex_wrap == "repeat" else GL_CLAMP
        gl.glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, wrap)
        gl.glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, wrap)
        t = np.ascontiguousarray(scene.texture, np.uint8)
        gl.glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA, t.shape[1], t.shape[0], 0, GL_RGBA, GL_UNSIGNED_BYTE, _ptr(t))
        loc = gl.glGetUniformLocation(prog, b"uTex")
        if loc >= 0:
            gl.glUniform1i(loc, 0)
    if scene.matrix is not None:
        loc = gl.glGetUniformLocation(prog, b"uM")
        if loc >= 0:
            m = np.ascontiguousarray(scene.matrix, np.float32)
            gl.glUniformMatrix4fv(loc, 1, GL_FALSE, m.ctypes.data_as(C.POINTER(_f32)))

    for name, (kind, vals) in getattr(scene, "uniforms", {}).items():
        loc = gl.glGetUniformLocation(prog, name.encode())
        if loc < 0:
            continue
        if kind == "1i":
            gl.glUniform1i(loc, int(vals[0]))
        elif kind in ("1f", "2f", "3f", "4f"):
            getattr(gl, "glUniform" + kind)(loc, *[float(v) for v in vals])
        else:
            m = np.ascontiguousarray(vals, np.float32)
            getattr(gl, f"glUniformMatrix{kind[1]}fv")(loc, 1, GL_FALSE, m.ctypes.data_as(C.POINTER(_f32)))

    vp = scene.viewport or (0, 0, scene.width, scene.height)
    gl.glViewport(*vp)
    gl.glClearColor(*scene.clear_color)
    return {"program": prog, "n_draw": n_draw, "indexed": indexed and scene.indices is not None,
            "vao": vao, "vbo": vbo, "ebo": ebo}


def frame_color(gl: GLApi, width: int, height: int) -> np.ndarray:
    """Copy of the colour attachment as uint32 [H, W] (row 0 = top, word = R<<24|G<<16|B<<8|A)."""
    p = gl.glGetFramePtr()
    return np.ctypeslib.as_array(p, shape=(height, width)).copy()


def fnv1a64(words: np.ndarray) -> int:
    """The survey's hash (appendix C): h = (h ^ word) * 1099511628211, seed 1469598103934665603.

    Evaluated blockwise in pure numpy-free Python for small arrays only (golden KATs)."""
    h = 1469598103934665603
    mask = (1 << 64) - 1
    for wv in np.asarray(words, dtype=np.uint32).reshape(-1).tolist():
        h = ((h ^ wv) * 1099511628211) & mask
    return h
