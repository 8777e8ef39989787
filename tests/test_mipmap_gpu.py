"""GPU: mip maps with a DEFINED level of detail.

The reference's per-triangle level goes through rsqrt(), whose `long` pun of a float is undefined
behaviour (swgl.c:3240-3246); the compiled reference never leaves the base level, which is the
library's default (tests/test_next_rows_gpu.py).  With swglSetOption("mip_lod", 1) the library
samples the glGenerateMipmap chain (k_mipmap_box, swgl.c:2122-2173) with the level the same code
gives when the pun is 32 bits wide; the checker is the reference compiled exactly that way
(oracle/ref_shim.c -DSWGLREF_DEFINED_RSQRT: `#define long int32_t` around the include, no source
edit).  north_star's colour tolerance is 1/255 per channel; every operation involved is an IEEE
binary32 operation in the reference's order, so the frames are expected to be identical and the
mismatch counters are asserted to be zero (they are reported on failure).
"""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import gl as G, scenes as S

pytestmark = pytest.mark.gpu

W, H = 320, 240

FS_TWO_TEXTURES = ("in vec4 vCol;\nuniform sampler2D texA;\nuniform sampler2D texB;\nout vec4 FragColor;\nvoid main()\n{\n"
                   "vec4 a = texture(texA,vCol.xy);\nvec4 b = texture(texB,vCol.zy);\nFragColor = a * vec4(0.5, 0.5, 0.5, 0.5) + b * vec4(0.5, 0.5, 0.5, 0.5);\n}\n")


@pytest.fixture(scope="module")
def reference_lod():
    try:
        return O.Reference(defined_rsqrt=True)
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/libswgl_ref_lod.so not available")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _frames(gpu_api, ref, script, path):
    out = []
    for api, is_ref in ((gpu_api, False), (ref.api, True)):
        api.glInit(W, H)
        if is_ref:
            ref.lib.swglref_fill(0x01020304, C.c_float(0.0))
        else:
            api.swglFillFramebuffer(0x01020304, C.c_float(0.0))
            api.swglSetOption(b"mip_lod", 1)
            api.swglSetOption(b"raster_path", path)
        api.glViewport(0, 0, W, H)
        api.glClearColor(0.0, 0.0, 0.0, 1.0)
        script(api)
        col = G.frame_color(api, W, H)
        if is_ref:
            dep = np.ctypeslib.as_array(ref.lib.swglref_depth_ptr(), shape=(H, W)).copy()
        else:
            dep = np.ctypeslib.as_array(api.swglGetDepthPtr(), shape=(H, W)).copy()
            assert api.swglGetLastError().decode() == ""
        out.append((col, dep))
    return out


def _check(a, b):
    cmp = O.compare(a[0], a[1], b[0], b[1])
    assert cmp["coverage_mismatch"] == 0 and cmp["depth_mismatch"] == 0, cmp
    assert cmp["max_channel_delta"] <= 1, f"colour off by more than 1/255: {cmp}"     # north_star tolerance
    assert cmp["color_mismatch"] == 0, f"within 1/255 but not identical: {cmp}"
    return cmp


@pytest.mark.parametrize("path", [1, 2, 3], ids=["pixel_owner", "fragment_parallel", "warp_tile"])
@pytest.mark.parametrize("grid,tex", [(24, "lcg1024"), (9, "lcg1024"), (60, "odd")])
def test_mip_chain_with_defined_lod_matches_the_reference_built_the_same_way(gpu_api, reference_lod, grid, tex, path):
    """Texture-shaped shader (the LOD moves it onto the generic fragment path): small, medium and large
    triangles pick different levels; the odd-sized float texture exercises the reference's 2*CurWidth
    row stride in the box filter and CLAMP addressing of the levels."""
    scene = S.grid_mesh(grid, W, H, textured=True)
    rng = np.random.default_rng(5)
    odd = rng.uniform(0.0, 1.0, (77, 93, 4)).astype(np.float32)   # RGBA: an RGB texture reads alpha 0 and blends to nothing
    frames = {}
    for mip in (False, True):
        def script(api, mip=mip):
            st = G.setup_scene(api, scene, indexed=False, init=False)
            if tex == "odd":
                t = C.c_uint32(0)
                api.glGenTextures(1, C.byref(t))
                api.glBindTexture(G.GL_TEXTURE_2D, t.value)
                api.glTexParameteri(G.GL_TEXTURE_2D, G.GL_TEXTURE_WRAP_S, G.GL_CLAMP)
                api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 93, 77, 0, G.GL_RGBA, G.GL_FLOAT, _ptr(odd))
            if mip:
                api.glGenerateMipmap(G.GL_TEXTURE_2D)
            api.glClear(3)
            api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
        a, b = _frames(gpu_api, reference_lod, script, path)
        _check(a, b)
        frames[mip] = b[0]
    # the chain is really sampled: the reference's own frame changes with it
    assert int((frames[False] != frames[True]).sum()) > 1000


def test_two_mipped_textures_in_a_generic_shader_and_a_near_clipped_scene(gpu_api, reference_lod):
    """The level comes from the CLIPPED triangle (swgl.c:3316 runs per DrawTriangle call); one texture has
    a chain, the other does not (base level whatever the LOD)."""
    scene = S.random_triangles(300, W, H, seed=77, near_cross=True)
    ta = S.lcg_texture(64, seed=3)
    tb = S.checker_texture(32)

    def script(api):
        v = api.glCreateShader(G.GL_VERTEX_SHADER)
        api.glShaderSource(v, S.VS_PASSTHROUGH.encode())
        api.glCompileShader(v)
        f = api.glCreateShader(G.GL_FRAGMENT_SHADER)
        api.glShaderSource(f, FS_TWO_TEXTURES.encode())
        api.glCompileShader(f)
        p = api.glCreateProgram()
        api.glAttachShader(p, v)
        api.glAttachShader(p, f)
        api.glLinkProgram(p)
        api.glUseProgram(p)
        verts = np.ascontiguousarray(scene.vertices, np.float32)
        vao, vbo = C.c_uint32(0), C.c_uint32(0)
        api.glGenVertexArrays(1, C.byref(vao))
        api.glBindVertexArray(vao.value)
        api.glGenBuffers(1, C.byref(vbo))
        api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
        api.glBufferData(G.GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), G.GL_STATIC_DRAW)
        api.glVertexAttribPointer(0, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(0))
        api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(16))
        t1, t2 = C.c_uint32(0), C.c_uint32(0)
        api.glGenTextures(1, C.byref(t1))
        api.glGenTextures(1, C.byref(t2))
        api.glActiveTexture(G.GL_TEXTURE1)
        api.glBindTexture(G.GL_TEXTURE_2D, t1.value)
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 64, 64, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(ta))
        api.glGenerateMipmap(G.GL_TEXTURE_2D)
        api.glActiveTexture(G.GL_TEXTURE2)
        api.glBindTexture(G.GL_TEXTURE_2D, t2.value)
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 32, 32, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(tb))
        api.glUniform1i(api.glGetUniformLocation(p, b"texA"), 1)
        api.glUniform1i(api.glGetUniformLocation(p, b"texB"), 2)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(verts))
    a, b = _frames(gpu_api, reference_lod, script, 3)
    _check(a, b)


def test_levels_outlive_their_image_and_second_chains_are_appended(gpu_api, reference_lod):
    """The reference never clears a texture's level vector (swgl.c:2094-2118, 2134-2166): glTexImage2D leaves the old
    image's levels in place (sampled with their own sizes), a second glGenerateMipmap pushes the new image's levels
    behind them, and min(level, count - 1) then reaches into the second chain.  Four frames, each after one more call."""
    scene = S.random_triangles(800, W, H, seed=5, extent=0.12, textured=True)      # levels from below 1 to beyond 10
    first = S.lcg_texture(64, seed=3)
    second = S.lcg_texture(16, seed=9)
    third = np.ascontiguousarray(S.lcg_texture(32, seed=11)[:, :, :3])       # fewer floats per texel than the levels: still inside them
    frames = []

    def script(api):
        st = G.setup_scene(api, scene, indexed=False, init=False)

        def frame():
            api.glClear(3)
            api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
            return G.frame_color(api, W, H).copy()
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 64, 64, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(first))
        api.glGenerateMipmap(G.GL_TEXTURE_2D)
        out = [frame()]
        api.glGenerateMipmap(G.GL_TEXTURE_2D)                    # the same four levels once more, behind the first four
        out.append(frame())
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 16, 16, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(second))
        out.append(frame())                                      # a 16x16 image with the 64x64 image's eight levels
        api.glGenerateMipmap(G.GL_TEXTURE_2D)                    # + two levels of the 16x16 image
        out.append(frame())
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGB, 32, 32, 0, G.GL_RGB, G.GL_UNSIGNED_BYTE, _ptr(third))
        out.append(frame())                                      # RGB image over RGBA levels: stride 3 inside 4-float texels
        frames.append(out)
    a, b = _frames(gpu_api, reference_lod, script, 3)
    _check(a, b)
    ours, theirs = frames
    for k, (x, y) in enumerate(zip(ours, theirs)):
        assert np.array_equal(x, y), f"frame {k}: {int((x != y).sum())} pixels differ"
    # the second chain, the third chain and the narrower image change what is sampled; replacing the image alone does
    # not change a pixel (a positive level only ever reads the old image's levels)
    changed = [int((theirs[k] != theirs[k + 1]).sum()) for k in range(4)]
    assert changed[0] > 500 and changed[1] == 0 and changed[2] > 500 and changed[3] > 500, changed


FS_POINT_TEX = ("in vec4 vCol;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\n"
                "vec4 t = texture(uTex,vCol.xy);\nFragColor = vec4(t.x, t.y, t.z, 1.0);\n}\n")


def _points_after_triangles(api, tri_scene, pts, with_triangles):
    v = api.glCreateShader(G.GL_VERTEX_SHADER); api.glShaderSource(v, S.VS_PASSTHROUGH.encode()); api.glCompileShader(v)
    f = api.glCreateShader(G.GL_FRAGMENT_SHADER); api.glShaderSource(f, FS_POINT_TEX.encode()); api.glCompileShader(f)
    p = api.glCreateProgram(); api.glAttachShader(p, v); api.glAttachShader(p, f); api.glLinkProgram(p); api.glUseProgram(p)
    verts = np.ascontiguousarray(np.concatenate([tri_scene.vertices, pts]), np.float32)
    vao, vbo, t = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    api.glGenVertexArrays(1, C.byref(vao)); api.glBindVertexArray(vao.value)
    api.glGenBuffers(1, C.byref(vbo)); api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
    api.glBufferData(G.GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), G.GL_STATIC_DRAW)
    api.glVertexAttribPointer(0, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(0))
    api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(16))
    tex = S.lcg_texture(64, seed=3)
    api.glGenTextures(1, C.byref(t)); api.glBindTexture(G.GL_TEXTURE_2D, t.value)
    api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 64, 64, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(tex))
    api.glUniform1i(api.glGetUniformLocation(p, b"uTex"), 0)
    api.glClear(3)
    nt = len(tri_scene.vertices)
    # one triangle first in both runs: the reference only allocates a program's fragment inputs when a triangle is shaded
    api.glDrawArrays(G.GL_TRIANGLES, 0, 3)
    api.glGenerateMipmap(G.GL_TEXTURE_2D)
    if with_triangles:
        api.glDrawArrays(G.GL_TRIANGLES, 3, nt - 3)          # the LAST of them leaves its level behind
    api.glClear(3)
    api.glDrawArrays(G.GL_POINTS, nt, len(pts))


def test_points_sample_with_the_level_the_last_triangle_left(gpu_api, reference_lod):
    """MipMapLevel is a global in the reference, set by every DrawTriangle call (swgl.c:3314-3316) and read by texture()
    whatever the primitive: a GL_POINTS draw samples with what the last triangle before it left (k_last_level)."""
    tri_scene = S.random_triangles(60, W, H, seed=8, extent=0.05, near_cross=True)        # small triangles: levels well above 1
    # ... behind one big triangle (level below 1: the half-size level), which is all the second run draws before its points
    tri_scene.vertices[0:3, 0:4] = [[-0.9, -0.9, 0.5, 1.0], [0.9, -0.8, 0.5, 1.0], [0.0, 0.9, 0.5, 1.0]]
    rng = np.random.default_rng(4)
    pts = np.zeros((3000, 8), np.float32)
    pts[:, 0:2] = rng.uniform(-0.95, 0.95, (3000, 2)); pts[:, 2] = 0.5; pts[:, 3] = 1.0
    pts[:, 4:8] = rng.uniform(0.0, 1.0, (3000, 4))
    frames = {}
    for with_triangles in (True, False):
        a, b = _frames(gpu_api, reference_lod, lambda api: _points_after_triangles(api, tri_scene, pts, with_triangles), 3)
        _check(a, b)
        frames[with_triangles] = b[0]
    # the level really reaches the points: without the triangles in between they sample another level
    assert int((frames[True] != frames[False]).sum()) > 300


@pytest.mark.parametrize("path", [3, 2], ids=["warp_tile", "fragment_parallel"])
def test_mip_levels_match_the_restatement_and_its_golden_frames(gpu_api, restatement, path):
    """The same frames against the C restatement of the mip chain and the defined level of detail (pinned on the CPU
    against the reference built that way, tests/test_oracle.py) and against tests/golden/lod_kats.json: colour, depth
    and the fragment counters, indexed draws included."""
    import json
    import os
    from test_oracle import lod_scenes
    from util import assert_bit_exact, gpu_render
    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lod_kats.json")))
    for name, scene in lod_scenes().items():
        rc, rd, rstats = restatement.render(scene, mipmaps=True)
        col, dep, stats, err = gpu_render(gpu_api, scene, indexed=scene.indices is not None, mipmaps=True,
                                          options={"mip_lod": 1, "raster_path": path})
        assert err == "", err
        assert_bit_exact(O.compare(col, dep, rc, rd), name)
        assert (stats["tested"], stats["shaded"]) == (rstats["tested"], rstats["shaded"])
        assert f"{restatement.fnv(col):016x}" == kats[name]["color_fnv"] and f"{restatement.fnv(dep):016x}" == kats[name]["depth_fnv"]


def test_default_stays_bug_compatible(gpu_api, reference):
    """Without the option the chain is built but never sampled, like the compiled reference."""
    scene = S.grid_mesh(24, W, H, textured=True)

    def run(api, is_ref):
        api.glInit(W, H)
        if is_ref:
            reference.lib.swglref_fill(0, C.c_float(0.0))
        else:
            api.swglFillFramebuffer(0, C.c_float(0.0))
        st = G.setup_scene(api, scene, indexed=False, init=False)
        api.glGenerateMipmap(G.GL_TEXTURE_2D)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
        return G.frame_color(api, W, H).copy()
    assert np.array_equal(run(gpu_api, False), run(reference.api, True))
    assert gpu_api.swglGetOption(b"mip_lod") == 0
