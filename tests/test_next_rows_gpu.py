"""GPU: the "next" rows of SURVEY.md 8(f) that are built -- GL_POINTS (n3) and glGenerateMipmap
(n2, bug-compatible with the compiled reference) -- against the compiled reference."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import gl as G, scenes as S

from test_api_edge_gpu import _assert_same, _both, _program, _ptr, _vao, W, H

pytestmark = pytest.mark.gpu


def _points(seed, n=6000):
    rng = np.random.default_rng(seed)
    v = np.empty((n, 8), np.float32)
    v[:, 3] = rng.uniform(0.5, 2.0, n)
    v[:, 0:2] = rng.uniform(-1.3, 1.3, (n, 2)) * v[:, 3:4]
    v[:, 2] = rng.uniform(-1.0, 1.0, n)
    v[:, 4:8] = rng.uniform(-0.2, 1.2, (n, 4))
    v[100:140] = v[100]                 # forty points on one pixel ...
    v[100:140, 4:8] = rng.uniform(0, 1, (40, 4))   # ... with different colours: the last one wins
    v[200, 3] = 0.0                     # w = 0
    return v


@pytest.mark.parametrize("viewport", [(0, 0, W, H), (13, 7, 150, 90)])
def test_points_pass_through(gpu_api, reference, viewport):
    v = _points(1)

    def script(api):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, v, [(0, 4, 0), (1, 4, 16)])
        api.glViewport(*viewport)
        api.glClear(3)
        # (the reference memcpy's into the fragment `in` variables before anything has allocated them
        # when GL_POINTS is a program's first draw, swgl.c:3548-3551; one triangle first allocates them)
        api.glDrawArrays(G.GL_TRIANGLES, 0, 3)
        api.glDrawArrays(G.GL_POINTS, 0, len(v))
        api.glDrawArrays(G.GL_POINTS, 50, 200)       # again: overwrites
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 1000)


def test_points_generic_shaders_and_texture_then_triangles(gpu_api, reference):
    v = _points(2)
    # a triangle that certainly produces fragments (on the near plane, in front of everything else):
    # the reference only allocates a program's fragment `in` variables when a fragment is shaded
    v[3597:3600, 0:4] = [[-0.5, -0.5, -1.0, 1.0], [0.5, -0.5, -1.0, 1.0], [0.0, 0.5, -1.0, 1.0]]
    vs = (S.VS_PASSTHROUGH.replace("gl_Position = aPos;", "gl_Position = aPos * vec4(0.9, 1.1, 1.0, 1.0);"))
    fs = ("in vec4 vCol;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\n"
          "vec4 t = texture(uTex,vCol.xy);\nFragColor = t * vCol.wzyx;\n}\n")
    tex = S.checker_texture(32)

    def script(api):
        p, _, _ = _program(api, vs, fs)
        api.glUseProgram(p)
        _vao(api, v, [(0, 4, 0), (1, 4, 16)])
        t = C.c_uint32(0)
        api.glGenTextures(1, C.byref(t))
        api.glActiveTexture(G.GL_TEXTURE0)
        api.glBindTexture(G.GL_TEXTURE_2D, t.value)
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 32, 32, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(tex))
        api.glUniform1i(api.glGetUniformLocation(p, b"uTex"), 0)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, 300)     # triangles, then points over them, then triangles again
        api.glDrawArrays(G.GL_POINTS, 300, 3000)
        api.glDrawArrays(G.GL_TRIANGLES, 3300, 300)
        p2, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_TEX_SWZ)
        api.glUseProgram(p2)
        api.glUniform1i(api.glGetUniformLocation(p2, b"uTex"), 0)
        api.glDrawArrays(G.GL_TRIANGLES, 3597, 3)    # allocates the new program's `in` variables in the reference
        api.glDrawArrays(G.GL_POINTS, 3600, 2000)
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 1000)


def test_generate_mipmap_keeps_sampling_the_base_level(gpu_api, reference):
    """In the compiled reference MipMapLevel is never positive (rsqrt's 8-byte pun, swgl.c:3246, makes
    it negative for every input), so a mip chain changes nothing; neither does it here."""
    scene = S.grid_mesh(24, W, H, textured=True)
    outs = {}
    for mip in (False, True):
        def script(api, mip=mip):
            st = G.setup_scene(api, scene, indexed=False, init=False)
            if mip:
                api.glGenerateMipmap(G.GL_TEXTURE_2D)
            api.glClear(3)
            api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
        a, b = _both(gpu_api, reference, script)
        _assert_same(a, b, 1000)
        outs[mip] = a
    assert np.array_equal(outs[False][0], outs[True][0])


def test_write_through_host_mirror_matches_copy_mode(gpu_api, reference):
    """n4: an application that reads every frame back gets the frame stored into the pinned mirror by
    the raster kernels themselves; what glGetFramePtr returns must not depend on the mode."""
    scene = S.random_triangles(400, W, H, seed=77, alpha=None, extent=0.4)
    pts = _points(5, 600)

    def frames(api, is_gpu):
        out = []
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, np.concatenate([scene.vertices, pts]), [(0, 4, 0), (1, 4, 16)])
        n = len(scene.vertices)
        for f in range(6):
            if f != 3:
                api.glClear(3)                     # frame 3 draws over frame 2 without a clear
            api.glDrawArrays(G.GL_TRIANGLES, (f * 30) % 300, n - 300)
            if f == 4:
                api.glDrawArrays(G.GL_POINTS, n, len(pts))     # a kernel that does not write through
            if f == 5:
                api.glDrawArrays(G.GL_TRIANGLES, 0, 90)        # second draw of the frame
            out.append(G.frame_color(api, W, H).copy())
        return out

    got = {}
    for mode in (1, 0):
        gpu_api.glInit(W, H)
        gpu_api.swglSetOption(b"host_mirror", mode)
        gpu_api.glViewport(0, 0, W, H)
        gpu_api.glClearColor(0.0, 0.0, 0.0, 1.0)
        w0 = gpu_api.swglGetOption(b"wt_draws")
        got[mode] = frames(gpu_api, True)
        used = gpu_api.swglGetOption(b"wt_draws") - w0
        assert (used > 0) == (mode == 1), used
        assert gpu_api.swglGetLastError().decode() == ""
    gpu_api.swglSetOption(b"host_mirror", 1)
    api = reference.api
    api.glInit(W, H)
    api.glViewport(0, 0, W, H)
    api.glClearColor(0.0, 0.0, 0.0, 1.0)
    want = frames(api, False)
    for f in range(6):
        assert np.array_equal(got[1][f], got[0][f]), f
        assert np.array_equal(got[1][f], want[f]), f
