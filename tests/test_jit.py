"""CPU: run-time shader compilation (swgl_jit.cpp) without a device.

swglPrecompileProgram() generates CUDA C++ from the bound program's op lists and compiles
k_vertex<SWVS_JIT> / k_raster_warp<SWFS_JIT> with NVRTC for sm_100a; with no GPU the cubin is not
loaded, but it can be inspected: the point of compiling instead of interpreting is
that the shader's variables live in registers, so the kernels must have no local-memory variable file
(the interpreter's raster kernel carries ~2 KB of stack and several hundred LDL/STL)."""
import ctypes as C
import glob
import os
import re
import subprocess

import pytest

import swgl_b200
from swgl_b200 import gl as G

from shader_cases import CASES


def _compile(api, vs, fs):
    v = api.glCreateShader(G.GL_VERTEX_SHADER)
    api.glShaderSource(v, vs.encode())
    api.glCompileShader(v)
    f = api.glCreateShader(G.GL_FRAGMENT_SHADER)
    api.glShaderSource(f, fs.encode())
    api.glCompileShader(f)
    p = api.glCreateProgram()
    api.glAttachShader(p, v)
    api.glAttachShader(p, f)
    api.glLinkProgram(p)
    api.glUseProgram(p)
    vao = C.c_uint32(0)
    api.glGenVertexArrays(1, C.byref(vao))
    api.glBindVertexArray(vao.value)
    api.glVertexAttribPointer(0, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(0))
    api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(16))
    return api.swglPrecompileProgram()


@pytest.mark.parametrize("name", sorted(CASES))
def test_every_generic_shader_compiles_to_register_resident_kernels(name, tmp_path, monkeypatch):
    monkeypatch.setenv("SWGL_JIT_DUMP", str(tmp_path))
    api = swgl_b200.load()
    api.swglGetLastError()
    vs, fs, _ = CASES[name]
    assert _compile(api, vs, fs) == 0, api.swglGetLastError().decode()
    cubins = glob.glob(os.path.join(str(tmp_path), "*.cubin"))
    if not cubins:
        pytest.skip("the program was compiled earlier in this process (cached)")
    usage = subprocess.run(["cuobjdump", "-res-usage", cubins[0]], capture_output=True, text=True).stdout
    kernels = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    assert kernels, usage
    for fn, reg, stack, shared, local in kernels:
        assert "ILi3E" in fn                                   # k_vertex<SWVS_JIT> / k_raster_warp<SWFS_JIT, 3>
        assert int(local) == 0 and int(stack) <= 128, (fn, stack, local)   # interpreter: STACK 1536 / 2096
        if "k_raster_warp" in fn:
            assert int(reg) <= 64                              # 8 CTAs of 128 threads per SM, like the built-in shapes
    sass = subprocess.run(["cuobjdump", "-sass", cubins[0]], capture_output=True, text=True).stdout
    assert len(re.findall(r"\b(LDL|STL)\b", sass)) <= 96       # a few caller-saved registers around cold calls; interpreter: 472
    src = open(glob.glob(os.path.join(str(tmp_path), "*.cu"))[0]).read()
    assert "jit_fragment" in src and "ir_execute(" not in src.split("generated from the program's IR")[1]


def test_built_in_shapes_need_no_compilation():
    from swgl_b200 import scenes as S
    api = swgl_b200.load()
    before = api.swglGetOption(b"jit_compiles")
    assert before >= 0
    assert _compile(api, S.VS_PASSTHROUGH, S.FS_COLOR) == 0
    assert api.swglGetOption(b"jit_compiles") == before
    vs, fs, _ = CASES["trig_parabola"]
    assert _compile(api, vs, fs) == 0
    n = api.swglGetOption(b"jit_compiles") + api.swglGetOption(b"jit_cache_hits")
    assert _compile(api, vs, fs) == 0                       # the same program again: served from the cache
    assert api.swglGetOption(b"jit_compiles") + api.swglGetOption(b"jit_cache_hits") == n + 1


def test_unlinked_program_reports_failure():
    api = swgl_b200.load()
    api.glUseProgram(0)
    api.glBindVertexArray(0)
    assert api.swglPrecompileProgram() == -1
