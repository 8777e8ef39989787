"""CPU: the quirk-compatible GLSL-subset front-end (swgl_glsl.c) -- IR shape per construct.

Each case cites the reference behaviour it pins."""
import ctypes as C
import struct

import pytest

import swgl_b200
from swgl_b200 import gl as G, scenes as S


@pytest.fixture(scope="module")
def api():
    return swgl_b200.load()


def compile_dump(api, kind, src):
    s = api.glCreateShader(kind)
    api.glShaderSource(s, src.encode())
    api.glCompileShader(s)
    buf = C.create_string_buffer(1 << 16)
    api.swglDebugShaderIR(s, buf, len(buf))
    api.swglGetLastError()
    return api.swglGetShaderCompiled(s), buf.value.decode()


def ops(dump):
    return [l.split()[2] for l in dump.splitlines() if l.startswith("op ")]


def fs(body, decls="in vec4 vCol;\nout vec4 FragColor;\n"):
    return decls + "void main()\n{\n" + body + "\n}\n"


def test_benchmark_shapes_are_recognised(api):
    ok, d = compile_dump(api, G.GL_VERTEX_SHADER, S.VS_PASSTHROUGH)
    assert ok and "pos_kind=1" in d and "simple_copies=1" in d
    ok, d = compile_dump(api, G.GL_VERTEX_SHADER, S.VS_MATRIX)
    assert ok and "pos_kind=2" in d and ops(d) == ["ldm", "ldv", "mulmv", "stv", "ldv", "stv"]
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, S.FS_COLOR)
    assert ok and "fs_kind=1" in d
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, S.FS_TEX)
    assert ok and "fs_kind=2" in d
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, S.FS_TEX_SWZ)
    assert ok and "fs_kind=2" in d and "swz" in ops(d)


def test_no_operator_precedence(api):
    """a + b * c folds left to right: (a + b) * c (swgl.c:1372-1408)."""
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("float a = 2.0 + 3.0 * 5.0;"))
    assert ok and ops(d)[:5] == ["conf", "conf", "add", "conf", "mul"]


def test_minus_digit_is_a_sign(api):
    """`-` directly followed by a digit is a sign, not an operator (swgl.c:933)."""
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("float a = 1.0 - 2.0;\nfloat b = -3.5;"))
    assert ok
    o = ops(d)
    assert o[:3] == ["conf", "conf", "sub"]
    imm = [l for l in d.splitlines() if " conf " in l][-1]
    bits = int(imm.split("imm=")[1].split()[0], 16)
    assert struct.unpack("<f", struct.pack("<I", bits))[0] == -3.5
    # "vCol.x-1.0": no operator is found, and the variable swizzle scanner skips the characters
    # it does not know (swgl.c:1330-1358) -- the statement silently means `a = vCol.x`
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("float a = vCol.x-1.0;"))
    assert ok and ops(d)[:3] == ["ldv", "swz", "stv"] and "sub" not in ops(d)


def test_swizzle_detection_stops_at_blank(api):
    """(a+b).xy is swizzled, (a + b).xy silently is not (swgl.c:1079-1092)."""
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("vec4 t = (vCol+vCol).wzyx;"))
    assert ok and "swz" in ops(d)
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("vec4 t = (vCol + vCol).wzyx;"))
    assert ok and "swz" not in ops(d)


def test_layout_with_in_keyword_is_not_an_attribute(api):
    """`layout (location = 0) in vec4 aPos;` becomes a never-fed `in` variable (swgl.c:1759-1765)."""
    ok, d = compile_dump(api, G.GL_VERTEX_SHADER, "layout (location = 0) in vec4 aPos;\nvoid main()\n{\ngl_Position = aPos;\n}\n")
    assert ok and "vec4 aPos word=4 in" in d and "layout" not in d.split("aPos")[1].splitlines()[0]


def test_type_mismatch_is_a_silent_noop(api):
    """vec4 * float -> GLSL_UNKNOWN; the assignment is skipped (swgl.c:2407-2411, 1898-1901)."""
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("FragColor = vCol * 2.0;"))
    assert ok and "stv" not in ops(d) and "zero" in ops(d)
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("FragColor = vCol * vec4(2.0, 2.0, 2.0, 1.0);"))
    assert ok and ops(d)[-1] == "stv"


def test_constructors_need_one_scalar_per_component(api):
    ok, _ = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("FragColor = vec4(vCol.xyz, 1.0);"))
    assert not ok  # Args.Size != 4 -> NULL token (swgl.c:1188-1194)
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("FragColor = vec4(vCol.x, 1, 0, 1);"))
    assert ok
    cons = [l for l in d.splitlines() if " cons " in l][0]
    assert "imm2=0xe" in cons  # args 1..3 are int-typed: converted with (float)i (swgl.c:2822-2827)


def test_matrix_ops_and_builtins(api):
    src = ("layout (location = 0) vec4 aPos;\nuniform mat4 A;\nuniform mat4 B;\nuniform mat3 N;\nout vec4 v;\n"
           "void main()\n{\nmat4 M = A * B + A;\ngl_Position = M * aPos;\nvec3 n = N * aPos.xyz;\n"
           "v = vec4(sin(n.x), cos(n.y), min(n.z, 0.5), max(aPos.w, 1.0));\n}\n")
    ok, d = compile_dump(api, G.GL_VERTEX_SHADER, src)
    assert ok, d
    o = ops(d)
    for name in ("mulmm", "addm", "stm", "mulmv", "sin", "cos", "min", "max", "cons"):
        assert name in o, (name, d)


def test_unsupported_constructs_fail_loudly(api):
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("float a = vCol.x < vCol.y;"))
    assert not ok and "comparison" in d
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, fs("FragColor = nothere;"))
    assert not ok and "unknown identifier" in d
    ok, d = compile_dump(api, G.GL_FRAGMENT_SHADER, "out vec4 FragColor;\n")
    assert not ok and "main" in d


def test_literal_parsers_follow_the_reference(api):
    lib = api.lib
    lib.swgl_glsl_atof.restype = C.c_double
    lib.swgl_glsl_atof.argtypes = [C.c_char_p]
    lib.swgl_glsl_atoi.restype = C.c_int
    lib.swgl_glsl_atoi.argtypes = [C.c_char_p]

    def ref_atof(s):  # swgl.c:18-59
        r, f, seen, sign, i = 0.0, 1.0, False, 1.0, 0
        if s[0] == "-":
            sign, i = -1.0, 1
        for ch in s[i:]:
            if ch == ".":
                seen = True
                continue
            if not ch.isdigit():
                break
            if seen:
                f /= 10.0
                r += f * int(ch)
            else:
                r = r * 10.0 + int(ch)
        return r * sign

    for s in ["0.1", "0.3", "3.14159", "-2.5", "100.001", "1.0f", "12", "0.7071067811865476"]:
        assert lib.swgl_glsl_atof(s.encode()) == ref_atof(s)
    assert lib.swgl_glsl_atoi(b"  -42abc") == -42 and lib.swgl_glsl_atoi(b"+7") == 7
