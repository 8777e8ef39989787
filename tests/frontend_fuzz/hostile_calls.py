"""Test infrastructure (tests/test_host_sanitizers.py): tens of thousands of API calls with ids, enums, sizes, offsets and
pointers that make no sense -- and a well-formed little scene now and then, so that the nonsense meets real objects --
against the host layer built with AddressSanitizer / UBSan over the stub device (SWGL_B200_LIB).  The reference has no
error channel and crashes on most of this; the host layer is held to ignoring it.

    python hostile_calls.py <seed> <calls>"""
import ctypes as C
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import swgl_b200
from swgl_b200 import gl as G, scenes as S
from shader_cases import CASES
from shader_fuzz_gen import make
api = swgl_b200.load()
rnd = random.Random(int(sys.argv[1]))
N = int(sys.argv[2])
SRC = [src for vs, fs, _ in CASES.values() for src in (vs, fs)] + [x for k in range(8) for x in make(k)[:2]] + ["", "garbage", "void main()\n{\n}\n", "uniform vec4 t"]
def ident(): return rnd.choice([0, 1, 2, 3, 4, 5, 7, 100, 65535, 0x7fffffff, 0xffffffff, rnd.randrange(0, 12)])
def enum(): return rnd.choice([G.GL_VERTEX_SHADER, G.GL_FRAGMENT_SHADER, G.GL_ARRAY_BUFFER, G.GL_ELEMENT_ARRAY_BUFFER, G.GL_TEXTURE_2D, G.GL_FLOAT, G.GL_UNSIGNED_BYTE, G.GL_RGB, G.GL_RGBA,
                               G.GL_TRIANGLES, G.GL_POINTS, G.GL_REPEAT, G.GL_CLAMP, G.GL_TEXTURE0, G.GL_TEXTURE0 + 7, G.GL_TEXTURE0 + 9, 0, 99, 0xffffffff, G.GL_UNSIGNED_INT, G.GL_STATIC_DRAW])
def small(): return rnd.choice([0, 1, 2, 3, 4, 5, 8, 16, 31, 32, 33, 64, 100, 255, 1000])
def anyint(): return rnd.choice([0, 1, -1, 3, 7, 100, -100, 0x7fffffff, -0x80000000, 65536, rnd.randrange(-50, 500)])
bufs = [np.ascontiguousarray(np.random.default_rng(k).uniform(-2, 2, (rnd.choice([0, 1, 3, 30, 300]), 8)), np.float32) for k in range(5)]
idxs = [np.ascontiguousarray(np.random.default_rng(k).integers(0, rnd.choice([1, 10, 400, 1 << 31]), rnd.choice([0, 1, 3, 90])), np.uint32) for k in range(4)]
texs = [np.ascontiguousarray(np.random.default_rng(k).integers(0, 255, (h, w, c)), np.uint8) for k, (h, w, c) in enumerate([(1, 1, 4), (8, 8, 4), (5, 7, 3), (64, 64, 4), (2, 300, 3), (16, 16, 4)])]
ftex = np.ascontiguousarray(np.random.default_rng(9).uniform(0, 1, (9, 9, 4)), np.float32)
mats = np.ascontiguousarray(np.random.default_rng(7).uniform(-1, 1, 16), np.float32)
out = C.c_uint32(0)
def ptr(a): return a.ctypes.data_as(C.c_void_p) if a.size else None
calls = 0
for it in range(N):
    if it % 400 == 0:
        api.glInit(rnd.choice([1, 7, 64, 200, 333]), rnd.choice([1, 5, 48, 152, 211]))
    r = rnd.randrange(40)
    try:
        if r == 0: s = api.glCreateShader(enum()); api.glShaderSource(s, rnd.choice(SRC).encode("latin-1", "replace")); api.glCompileShader(s)
        elif r == 1: api.glShaderSource(ident(), rnd.choice(SRC).encode("latin-1", "replace"))
        elif r == 2: api.glCompileShader(ident())
        elif r == 3: api.glCreateProgram()
        elif r == 4: api.glAttachShader(ident(), ident())
        elif r == 5: api.glLinkProgram(ident())
        elif r == 6: api.glUseProgram(ident())
        elif r == 7: api.glGenVertexArrays(small(), C.byref(out)); api.glBindVertexArray(rnd.choice([out.value, ident()]))
        elif r == 8: api.glBindVertexArray(ident())
        elif r == 9: api.glGenBuffers(small(), C.byref(out)); api.glBindBuffer(enum(), rnd.choice([out.value, ident()]))
        elif r == 10: api.glBindBuffer(enum(), ident())
        elif r == 11: b = rnd.choice(bufs); api.glBufferData(enum(), rnd.choice([b.nbytes, 0, b.nbytes // 2]), ptr(b), enum())
        elif r == 12: b = rnd.choice(idxs); api.glBufferData(G.GL_ELEMENT_ARRAY_BUFFER, b.nbytes, ptr(b), G.GL_STATIC_DRAW)
        elif r == 13: api.glVertexAttribPointer(ident(), anyint(), enum(), 0, rnd.choice([0, 4, 32, 36, 1 << 20, 0xffffffff]), C.c_void_p(rnd.choice([0, 16, 28, 1 << 20, (1 << 40)])))
        elif r == 14: api.glEnableVertexAttribArray(ident())
        elif r == 15: api.glViewport(anyint(), anyint(), rnd.choice([0, 1, 64, 200, 5000, 0x7fffffff, 0xffffffff]), rnd.choice([0, 1, 48, 152, 5000, 0x7fffffff, 0xffffffff]))
        elif r == 16: api.glClearColor(rnd.uniform(-2, 2), rnd.uniform(-2, 2), float("nan") if rnd.random() < 0.1 else 0.5, float("inf") if rnd.random() < 0.1 else 1.0)
        elif r == 17: api.glClear(rnd.choice([0, 1, 2, 3, 7, 0xffffffff]))
        elif r == 18: api.glDrawArrays(enum(), anyint(), rnd.choice([0, 1, 2, 3, 30, 299, 300, 301, 10000, 0x7fffffff, 0xffffffff]))
        elif r == 19: api.glDrawElements(enum(), rnd.choice([0, 1, 3, 89, 90, 91, 100000, 0xffffffff]), enum(), C.c_void_p(rnd.choice([0, 4, 12, 1 << 20, 3])))
        elif r == 20: api.glGenTextures(small(), C.byref(out)); api.glBindTexture(enum(), rnd.choice([out.value, ident()]))
        elif r == 21: api.glActiveTexture(enum())
        elif r == 22: api.glBindTexture(enum(), ident())
        elif r == 23: api.glTexParameteri(enum(), rnd.choice([G.GL_TEXTURE_WRAP_S, G.GL_TEXTURE_WRAP_T, 0, 99]), enum())
        elif r == 24:
            t = rnd.choice(texs); fmt = G.GL_RGBA if t.shape[2] == 4 else G.GL_RGB
            api.glTexImage2D(rnd.choice([G.GL_TEXTURE_2D, enum()]), anyint() % 3, rnd.choice([fmt, enum()]), rnd.choice([t.shape[1], 0]), rnd.choice([t.shape[0], 0]), rnd.choice([0, 0, 1]), fmt, G.GL_UNSIGNED_BYTE, ptr(t))
        elif r == 25: api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 9, 9, 0, G.GL_RGBA, rnd.choice([G.GL_FLOAT, enum()]), ptr(ftex))
        elif r == 26: api.glGenerateMipmap(enum())
        elif r == 27: api.glGetUniformLocation(ident(), rnd.choice([b"tint", b"uM", b"uTex", b"", b"nope", b"u4", b"A", b"x" * 300]))
        elif r == 28:
            loc = rnd.choice([api.glGetUniformLocation(ident(), rnd.choice([b"tint", b"uM", b"uTex", b"u4", b"A", b"k", b"N", b"R"])), -1, 0, 1, 0x10000, 0x7fffffff, anyint()])
            k = rnd.randrange(8)
            if k == 0: api.glUniform1f(loc, 0.5)
            elif k == 1: api.glUniform2f(loc, 0.5, 0.25)
            elif k == 2: api.glUniform3f(loc, 0.5, 0.25, 1.0)
            elif k == 3: api.glUniform4f(loc, 0.5, 0.25, 1.0, 2.0)
            elif k == 4: api.glUniform1i(loc, anyint())
            elif k == 5: api.glUniformMatrix2fv(loc, rnd.choice([0, 1, 2]), rnd.choice([0, 1]), mats.ctypes.data_as(C.POINTER(C.c_float)))
            elif k == 6: api.glUniformMatrix3fv(loc, 1, rnd.choice([0, 1]), mats.ctypes.data_as(C.POINTER(C.c_float)))
            else: api.glUniformMatrix4fv(loc, 1, rnd.choice([0, 1]), mats.ctypes.data_as(C.POINTER(C.c_float)))
        elif r == 29: api.glGetFramePtr(); api.swglGetDepthPtr()
        elif r == 30: b = rnd.choice(bufs); api.swglBufferRespecify(enum(), rnd.choice([b.nbytes, 0, 12]), ptr(b))
        elif r == 31: b = rnd.choice(idxs); api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, b.nbytes, ptr(b))
        elif r == 32: b = rnd.choice(bufs); api.swglBufferSubData(enum(), rnd.choice([0, 4, 1 << 20, (1 << 40)]), rnd.choice([b.nbytes, 4, 0]), ptr(b))
        elif r == 33: api.swglPrecompileProgram()
        elif r == 34: buf = (C.c_char * rnd.choice([1, 16, 4096]))(); api.swglDebugShaderIR(ident(), buf, len(buf)); api.swglGetShaderCompiled(ident())
        elif r == 35: api.glDeleteShader(ident())
        elif r == 36: api.swglFrameWait(rnd.choice([0, 1, 5, api.swglFrameSubmit()]))
        elif r == 37: api.swglSetOption(rnd.choice([b"jit", b"mip_lod", b"nope", b"bin_cap", b""]), anyint()); api.swglGetOption(rnd.choice([b"jit_compiles", b"device_count", b"nope"]))
        elif r == 38: api.swglGetLastError(); st = swgl_b200.swglStats(); api.swglGetStats(C.byref(st))
        else:
            # a well-formed little scene now and then, so that later garbage meets real objects
            sc = S.random_triangles(20, 64, 48, seed=it)
            G.setup_scene(api, sc, indexed=False, init=False)
            api.glDrawArrays(G.GL_TRIANGLES, 0, 60)
        calls += 1
    except C.ArgumentError:
        pass
print("hostile calls:", calls, flush=True)
os._exit(0)
