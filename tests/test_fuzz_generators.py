"""CPU: the random generators behind the GPU fuzz tests produce what they promise -- sequences and shaders inside the
reference's DEFINED behaviour.  The compiled reference runs every generated call sequence here (a sequence outside its
defined behaviour shows as a crash or as a frame that changes between two runs), the library's front end accepts
every generated shader, and a few of them go through the run-time compiler (no device needed)."""
import ctypes as C

import numpy as np
import pytest

import swgl_b200
from oracle import pyoracle as O
from swgl_b200 import gl as G

import test_api_fuzz_gpu as F
from shader_fuzz_gen import make
from test_jit import _compile


def _run_on_reference(ref, ops, points=True):
    W, H = ops[0][3], ops[0][4]
    return F.run_ops(ref.api, ops, lambda w, d: ref.lib.swglref_fill(w, C.c_float(d)),
                     lambda: np.ctypeslib.as_array(ref.lib.swglref_depth_ptr(), shape=(H, W)).copy(), points=points)


def test_call_sequences_are_deterministic_on_the_compiled_reference(reference):
    kinds = set()
    for seed in range(0, 60):
        ops = F.make_ops(seed)
        kinds.update(o[0] for o in ops)
        f1, d1 = _run_on_reference(reference, ops)
        f2, d2 = _run_on_reference(reference, ops)
        assert len(f1) == len(f2) and all(np.array_equal(a, b) for a, b in zip(f1, f2)), seed
        assert np.array_equal(d1.view(np.uint32), d2.view(np.uint32)), seed
    # the generator reaches every kind of call it knows
    assert {"setup", "use", "vao", "draw", "points", "viewport", "clear", "clearcolor", "matrix", "tint", "uni", "sampler",
            "wrap", "teximage", "mipmap", "bindtex", "respecify", "read"} <= kinds
    assert F.make_ops(7)[1:5] == F.make_ops(7)[1:5]


def test_mip_level_sequences_run_on_the_defined_rsqrt_reference():
    try:
        ref = O.Reference(defined_rsqrt=True)
    except Exception as e:                                  # pragma: no cover
        pytest.skip(f"oracle/_ref/libswgl_ref_lod.so not available: {e}")
    for seed in range(1000, 1030):
        ops = F.make_ops(seed, lod=True)
        assert ops[1] == ("mipmap", 1)
        f1, _ = _run_on_reference(ref, ops)
        f2, _ = _run_on_reference(ref, ops)
        assert all(np.array_equal(a, b) for a, b in zip(f1, f2)), seed


def test_random_shaders_are_accepted_by_the_front_end_and_compile():
    api = swgl_b200.load()
    for seed in range(40):
        vs, fs, uniforms = make(seed)
        assert make(seed) == (vs, fs, uniforms)
        for kind, src in ((G.GL_VERTEX_SHADER, vs), (G.GL_FRAGMENT_SHADER, fs)):
            sh = api.glCreateShader(kind)
            api.glShaderSource(sh, src.encode())
            api.glCompileShader(sh)
            assert api.swglGetShaderCompiled(sh) == 1, (seed, src)
    api.swglGetLastError()
    for seed in (900, 903, 905):                            # three of the programs tests/test_api_fuzz_gpu.py pools
        vs, fs, _ = make(seed)
        assert _compile(api, vs, fs) == 0, api.swglGetLastError().decode()


def test_generator_modes_keep_their_promises():
    for seed in range(200):
        ops = F.make_ops(seed, inside=True)
        H = ops[0][4]
        assert not any(o[0] == "viewport" and (o[2] < 0 or o[2] + o[4] > H) for o in ops), seed
        # under `lod` an image never has more floats per texel than the levels its texture object already carries
        ops = F.make_ops(seed, lod=True)
        bound, cur, levels = 1, {0: 4, 1: 4}, {1: 4}
        for o in ops[2:]:
            if o[0] == "bindtex":
                bound = o[2]
            elif o[0] == "teximage":
                assert o[2].shape[2] <= levels.get(bound, 4), seed
                cur[bound] = o[2].shape[2]
            elif o[0] == "mipmap":
                levels[bound] = min(levels.get(bound, 4), cur[bound])
        # draws stay inside the array they read (a partial last triple is read whole)
        arrays = [a for a in F.make_ops(seed)[0][1]]
        vao = 0
        for o in F.make_ops(seed)[1:]:
            if o[0] == "vao":
                vao = o[1]
            elif o[0] == "respecify":
                vao, arrays[o[1]] = o[1], o[2]
            elif o[0] == "draw":
                n = len(arrays[vao][0]) if arrays[vao][1] is None else len(arrays[vao][1])
                assert o[1] + 3 * ((o[2] + 2) // 3) <= n, (seed, o, n)
            elif o[0] == "points":
                assert arrays[vao][1] is None and o[1] + o[2] <= len(arrays[vao][0]), (seed, o)


def test_object_ids_and_uniform_locations_match_the_compiled_reference(reference):
    """Every shader case and 25 random programs, created in the same order in both libraries: shader / program / vertex
    array / buffer / texture ids, the glGen* return values and the location of 25 names (uniforms of both stages, names
    that are not uniforms, names that do not exist) must be the same numbers (no device needed)."""
    from shader_cases import CASES
    names = [b"tint", b"k", b"freq", b"A", b"B", b"N", b"R", b"uTex", b"u1", b"u2", b"u3", b"u4", b"uM", b"steps", b"shift", b"gain",
             b"offs", b"nope", b"", b"vCol", b"aPos", b"gl_Position", b"FragColor", b"t0", b"res"]
    progs = [(vs, fs) for vs, fs, _ in CASES.values()] + [make(k)[:2] for k in range(25)]
    rows = []
    for api in (swgl_b200.load(), reference.api):
        api.glInit(32, 32)                       # (without a device the library's glInit only resets the object tables)
        out = []
        for vs, fs in progs:
            v = api.glCreateShader(G.GL_VERTEX_SHADER); api.glShaderSource(v, vs.encode()); api.glCompileShader(v)
            f = api.glCreateShader(G.GL_FRAGMENT_SHADER); api.glShaderSource(f, fs.encode()); api.glCompileShader(f)
            p = api.glCreateProgram(); api.glAttachShader(p, v); api.glAttachShader(p, f); api.glLinkProgram(p)
            vao, vbo, tex = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
            r1, r2 = api.glGenVertexArrays(1, C.byref(vao)), api.glGenBuffers(1, C.byref(vbo))
            api.glGenTextures(1, C.byref(tex))
            out.append((v, f, p, vao.value, vbo.value, tex.value, r1, r2, tuple(api.glGetUniformLocation(p, n) for n in names)))
        rows.append(out)
    swgl_b200.load().swglGetLastError()
    assert rows[0] == rows[1]
    assert any(loc >= 0 for row in rows[1] for loc in row[8]) and any(loc == -1 for row in rows[1] for loc in row[8])
