"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute)."""
import ctypes as C
import os
import re

import swgl_b200
from swgl_b200 import gl as G
from swgl_b200._lib import EXTENSION_EXPORTS, LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROTO = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\**\s+\**\s*((?:gl|swgl|swgldev_)[A-Za-z0-9_]+)\s*\(", re.M)


def declared(header):
    with open(os.path.join(ROOT, "include", header)) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(PROTO.findall(text)))


def test_library_exists_in_tree():
    assert os.path.exists(LIB_PATH), "build with python -m swgl_b200.build"


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(LIB_PATH)
    names = declared("swgl.h") + declared("swgl_b200.h") + declared("swgl_dev.h")
    assert len(names) > 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_api_is_complete():
    """All 36 reference entry points (swgl.h:87-158) plus the glDrawElements extension."""
    names = declared("swgl.h")
    assert set(G.REFERENCE_EXPORTS) <= set(names)
    assert len(G.REFERENCE_EXPORTS) == 36
    assert "glDrawElements" in names


def test_python_binding_covers_extensions():
    names = set(declared("swgl_b200.h"))
    assert names <= set(EXTENSION_EXPORTS) | set(G.REFERENCE_EXPORTS), names - set(EXTENSION_EXPORTS)


def test_enum_positions_match_reference_header():
    """Positional enum values (swgl.h:40-81) must be preserved; extensions are appended."""
    with open(os.path.join(ROOT, "include", "swgl.h")) as f:
        text = f.read()
    body = text[text.index("typedef enum"):text.index("} GLenum;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b(GL_[A-Z0-9_]+)\b", body)
    assert names.index("GL_VERTEX_SHADER") == 0
    assert names.index("GL_ARRAY_BUFFER") == 4 and names.index("GL_STATIC_DRAW") == 5
    assert names.index("GL_FLOAT") == 8 and names.index("GL_UNSIGNED_BYTE") == 10
    assert names.index("GL_RGB") == 15 and names.index("GL_RGBA") == 16 and names.index("GL_TRIANGLES") == 17
    assert names.index("GL_REPEAT") == 20 and names.index("GL_CLAMP") == 21 and names.index("GL_TEXTURE_2D") == 22
    assert names.index("GL_TEXTURE0") == 25 and names.index("GL_TEXTURE7") == 32
    assert names.index("GL_ELEMENT_ARRAY_BUFFER") == 33 and names.index("GL_UNSIGNED_INT") == 34
    for n in names:
        assert getattr(G, n) == names.index(n)


def test_object_id_conventions_without_device():
    """Shader ids start at 0, program ids at 1, uniform location = ((program-1)<<16)|index
    (swgl.c:2875, 2922, 3757).  Pure host state: runs without a GPU."""
    api = swgl_b200.load()
    s0 = api.glCreateShader(G.GL_VERTEX_SHADER)
    api.glShaderSource(s0, b"uniform mat4 uM;\nuniform float k;\nvoid main()\n{\ngl_Position = uM * gl_Position;\n}\n")
    api.glCompileShader(s0)
    s1 = api.glCreateShader(G.GL_FRAGMENT_SHADER)
    api.glShaderSource(s1, b"uniform float k;\nuniform vec4 tint;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = tint;\n}\n")
    api.glCompileShader(s1)
    assert s1 == s0 + 1
    assert api.swglGetShaderCompiled(s0) == 1 and api.swglGetShaderCompiled(s1) == 1
    p = api.glCreateProgram()
    assert p >= 1
    api.glAttachShader(p, s0)
    api.glAttachShader(p, s1)
    api.glLinkProgram(p)
    base = (p - 1) << 16
    assert api.glGetUniformLocation(p, b"uM") == base | 0
    assert api.glGetUniformLocation(p, b"k") == base | 1       # the vertex-stage copy wins (swgl.c:3749-3758)
    assert api.glGetUniformLocation(p, b"tint") == base | 3     # VS uniforms first, then FS
    assert api.glGetUniformLocation(p, b"nope") == -1
    api.glUniform1f(-1, 1.0)                                    # guarded; UB in the reference
    vao = C.c_uint32(0)
    assert api.glGenVertexArrays(1, C.byref(vao)) == 0 and vao.value >= 1
    buf = C.c_uint32(0)
    assert api.glGenBuffers(1, C.byref(buf)) == 0 and buf.value >= 1


def test_extensions_are_safe_without_a_device_context():
    """The round's extensions before glInit (or on a box without a GPU): no-ops with the documented
    failure values, never a crash -- and never a CPU fallback."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("the context-less behaviour is checked on the CPU box")
    api = swgl_b200.load()
    word = (C.c_uint32 * 4)(1, 2, 3, 4)
    api.swglBufferSubData(G.GL_ARRAY_BUFFER, 0, 16, C.cast(word, C.c_void_p))
    api.swglBufferDeviceWritten(G.GL_ELEMENT_ARRAY_BUFFER)
    assert api.swglGetBufferDevicePtr(G.GL_ARRAY_BUFFER) == 0
    assert api.swglSetSharedFrameMirror(C.cast(word, C.c_void_p), 16) == -1
    assert not api.swglHostAlloc(4096, 1)          # cudaHostAlloc needs a device: NULL, not malloc
    api.swglHostFree(None)
    assert not api.glGetFramePtr()
