"""GPU: random SEQUENCES of the reference's own API calls -- programs, vertex arrays, textures, uniforms, viewports
(also ones that leave the framebuffer), partial clears, draws with arbitrary first / count, frames read back in the
middle -- issued identically to libswgl_b200.so and to the compiled reference (oracle/_ref).  Every frame read on the
way and the final colour words and depth bits must be identical.

Only behaviour that is DEFINED in the reference is generated (no reads past a buffer, well-typed shaders, vec4
fragment output).  tools/api_fuzz.py runs many more seeds of the same generator."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import gl as G, scenes as S

pytestmark = pytest.mark.gpu

W, H = 200, 152

VS_COLOUR = S.VS_PASSTHROUGH
FS_COLOUR = S.FS_COLOR
VS_MATRIX = ("layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\nuniform mat4 uM;\nout vec4 vCol;\n"
             "void main()\n{\ngl_Position = uM * aPos;\nvCol = aCol;\n}\n")
VS_UV = ("layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\nout vec2 vUV;\n"
         "void main()\n{\ngl_Position = aPos;\nvUV = aCol.xy;\n}\n")
FS_TEX = "in vec2 vUV;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = texture(uTex, vUV);\n}\n"
FS_TINT = ("in vec4 vCol;\nuniform vec4 tint;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\n"
           "vec4 t = texture(uTex,vCol.zy) * tint;\nFragColor = vec4(t.x, vCol.y, t.z, vCol.w);\n}\n")
PROGRAMS = [(VS_COLOUR, FS_COLOUR), (VS_MATRIX, FS_COLOUR), (VS_UV, FS_TEX), (VS_COLOUR, FS_TINT), (VS_MATRIX, FS_TINT)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def make_ops(seed):
    """The call sequence of a seed as a list of tuples (data only: the same list drives both libraries)."""
    rng = np.random.default_rng(31000 + seed)
    arrays = []
    for k in range(int(rng.integers(2, 4))):
        sc = S.random_triangles(int(rng.integers(40, 260)), W, H, seed=int(rng.integers(1, 1 << 30)),
                                extent=float(rng.choice([0.08, 0.25, 0.6, 1.3])),
                                alpha=None if rng.random() < 0.5 else float(rng.choice([1.0, 0.5, 0.15])),
                                near_cross=bool(rng.random() < 0.4), centre_range=float(rng.choice([0.8, 1.0, 1.3])))
        arrays.append(np.ascontiguousarray(sc.vertices, np.float32))
    textures = [S.checker_texture(int(rng.choice([8, 32]))), S.lcg_texture(int(rng.choice([16, 64])), seed=int(rng.integers(1, 99)))]
    ops = [("setup", arrays, textures)]
    prog, vao = 0, 0
    ops += [("use", 0), ("vao", 0), ("clear", 3)]
    for _ in range(int(rng.integers(10, 26))):
        r = rng.random()
        if r < 0.38:
            n = len(arrays[vao])
            first = int(rng.integers(0, n - 3))
            if rng.random() < 0.8:
                first -= first % 3
            count = int(rng.integers(0, n - first + 1))
            if first + 3 * ((count + 2) // 3) > n:        # a partial last triple is read whole: keep it inside the buffer
                count = 3 * ((n - first) // 3)
            ops.append(("draw", first, count))
        elif r < 0.50:
            if rng.random() < 0.5:
                ops.append(("viewport", 0, 0, W, H))
            else:
                x = int(rng.integers(-24, W // 2)) if rng.random() < 0.3 else int(rng.integers(0, W // 2))
                y = int(rng.integers(-40, H // 2)) if rng.random() < 0.3 else int(rng.integers(0, H // 2))
                w = int(rng.integers(8, W - max(x, 0) + (24 if rng.random() < 0.3 else 0) + 1))
                h = int(rng.integers(8, H - max(y, 0) + (48 if rng.random() < 0.3 else 0) + 1))
                ops.append(("viewport", x, y, w, h))
        elif r < 0.60:
            ops.append(("clearcolor",) + tuple(float(v) for v in rng.uniform(-0.2, 1.2, 4)))
            ops.append(("clear", int(rng.integers(1, 4))))
        elif r < 0.70:
            prog = int(rng.integers(len(PROGRAMS)))
            ops.append(("use", prog))
        elif r < 0.78:
            vao = int(rng.integers(len(arrays)))
            ops.append(("vao", vao))
        elif r < 0.86:
            m = np.eye(4, dtype=np.float32) + rng.uniform(-0.15, 0.15, (4, 4)).astype(np.float32)
            ops.append(("matrix", m))
        elif r < 0.91:
            ops.append(("tint",) + tuple(float(v) for v in rng.uniform(0.1, 1.3, 4)))
        elif r < 0.95:
            ops.append(("sampler", int(rng.integers(0, 2))))
        elif r < 0.97:
            ops.append(("wrap", int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 2))))
        else:
            ops.append(("read",))
    return ops


def run_ops(api, ops, fill, depth_of):
    """Issue `ops`; -> (frames read on the way + the final one, final depth)."""
    api.glInit(W, H)
    fill(0x0A0B0C0D, 0.0)
    api.glViewport(0, 0, W, H)
    api.glClearColor(0.0, 0.0, 0.0, 1.0)
    frames, progs, vaos, cur = [], [], [], 0
    for op in ops:
        k = op[0]
        if k == "setup":
            for vs, fs in PROGRAMS:
                v = api.glCreateShader(G.GL_VERTEX_SHADER); api.glShaderSource(v, vs.encode()); api.glCompileShader(v)
                f = api.glCreateShader(G.GL_FRAGMENT_SHADER); api.glShaderSource(f, fs.encode()); api.glCompileShader(f)
                p = api.glCreateProgram(); api.glAttachShader(p, v); api.glAttachShader(p, f); api.glLinkProgram(p)
                progs.append(p)
                # every uniform gets a value before the first draw (the reference's storage starts uninitialised)
                api.glUseProgram(p)
                ident = np.eye(4, dtype=np.float32)
                for name, setter in ((b"uM", lambda l: api.glUniformMatrix4fv(l, 1, G.GL_FALSE, ident.ctypes.data_as(C.POINTER(C.c_float)))),
                                     (b"tint", lambda l: api.glUniform4f(l, 1.0, 1.0, 1.0, 1.0)), (b"uTex", lambda l: api.glUniform1i(l, 0))):
                    loc = api.glGetUniformLocation(p, name)
                    if loc >= 0:
                        setter(loc)
            for verts in op[1]:
                vao, vbo = C.c_uint32(0), C.c_uint32(0)
                api.glGenVertexArrays(1, C.byref(vao)); api.glBindVertexArray(vao.value)
                api.glGenBuffers(1, C.byref(vbo)); api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
                api.glBufferData(G.GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), G.GL_STATIC_DRAW)
                api.glVertexAttribPointer(0, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(0))
                api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(16))
                vaos.append(vao.value)
            for unit, tex in enumerate(op[2]):
                t = C.c_uint32(0)
                api.glGenTextures(1, C.byref(t))
                api.glActiveTexture(G.GL_TEXTURE0 + unit)
                api.glBindTexture(G.GL_TEXTURE_2D, t.value)
                tt = np.ascontiguousarray(tex, np.uint8)
                api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, tt.shape[1], tt.shape[0], 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(tt))
        elif k == "use":
            cur = progs[op[1]]
            api.glUseProgram(cur)
        elif k == "vao":
            api.glBindVertexArray(vaos[op[1]])
        elif k == "draw":
            api.glDrawArrays(G.GL_TRIANGLES, op[1], op[2])
        elif k == "viewport":
            api.glViewport(*op[1:])
        elif k == "clearcolor":
            api.glClearColor(*op[1:])
        elif k == "clear":
            api.glClear(op[1])
        elif k == "matrix":
            loc = api.glGetUniformLocation(cur, b"uM")
            if loc >= 0:
                api.glUniformMatrix4fv(loc, 1, G.GL_FALSE, op[1].ctypes.data_as(C.POINTER(C.c_float)))
        elif k == "tint":
            loc = api.glGetUniformLocation(cur, b"tint")
            if loc >= 0:
                api.glUniform4f(loc, *op[1:])
        elif k == "sampler":
            loc = api.glGetUniformLocation(cur, b"uTex")
            if loc >= 0:
                api.glUniform1i(loc, op[1])
        elif k == "wrap":
            api.glActiveTexture(G.GL_TEXTURE0 + op[1])
            api.glTexParameteri(G.GL_TEXTURE_2D, G.GL_TEXTURE_WRAP_S if op[2] else G.GL_TEXTURE_WRAP_T, G.GL_CLAMP if op[3] else G.GL_REPEAT)
        elif k == "read":
            frames.append(G.frame_color(api, W, H))
    frames.append(G.frame_color(api, W, H))
    return frames, depth_of()


def compare_seed(gpu_api, reference, seed):
    """-> '' if the two libraries agree on every frame of the sequence, else a description."""
    ops = make_ops(seed)
    gf, gd = run_ops(gpu_api, ops, lambda w, d: gpu_api.swglFillFramebuffer(w, C.c_float(d)),
                     lambda: np.ctypeslib.as_array(gpu_api.swglGetDepthPtr(), shape=(H, W)).copy())
    err = gpu_api.swglGetLastError().decode()
    if err:
        return f"seed {seed}: {err}"
    rf, rd = run_ops(reference.api, ops, lambda w, d: reference.lib.swglref_fill(w, C.c_float(d)),
                     lambda: np.ctypeslib.as_array(reference.lib.swglref_depth_ptr(), shape=(H, W)).copy())
    for i, (a, b) in enumerate(zip(gf, rf)):
        if not np.array_equal(a, b):
            return f"seed {seed}: frame {i} of {len(gf)} differs in {int((a != b).sum())} pixels; ops {[o[0] for o in ops]}"
    cmp = O.compare(gf[-1], gd, rf[-1], rd)
    if cmp["color_mismatch"] or cmp["depth_mismatch"] or cmp["coverage_mismatch"]:
        return f"seed {seed}: {cmp}"
    return ""


@pytest.mark.parametrize("seed", range(12))
def test_random_call_sequence_matches_compiled_reference(gpu_api, reference, seed):
    assert compare_seed(gpu_api, reference, seed) == ""
