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


def _two_buffer_sets(api, n_bytes_v, attribs):
    """Two VAOs with their own vertex buffers (the double-buffered geometry of a streaming client)."""
    sets = []
    for _ in range(2):
        # data specified with no vertex array bound, so that the NAMED buffer owns it and the set can
        # be bound again by name (with an array bound, glBindBuffer snapshots the name, swgl.c:3116-3122)
        api.glBindVertexArray(0)
        vbo = C.c_uint32(0)
        api.glGenBuffers(1, C.byref(vbo))
        api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
        zero = np.zeros(n_bytes_v // 4, np.float32)
        api.glBufferData(G.GL_ARRAY_BUFFER, n_bytes_v, _ptr(zero), G.GL_STATIC_DRAW)
        vao = C.c_uint32(0)
        api.glGenVertexArrays(1, C.byref(vao))
        api.glBindVertexArray(vao.value)
        api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
        for loc, n, off in attribs:
            api.glVertexAttribPointer(loc, n, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(off))
            api.glEnableVertexAttribArray(loc)
        sets.append((vao.value, vbo.value))
    return sets


@pytest.mark.parametrize("full_clear", [True, False])
def test_frame_pipelining_submit_wait(gpu_api, reference, full_clear):
    """n4: swglFrameSubmit / swglFrameWait over two mirrors, uploads overlapping the previous frame."""
    frames_v = [np.ascontiguousarray(S.random_triangles(300, W, H, seed=900 + f, alpha=None, extent=0.5).vertices)
                for f in range(7)]
    nbytes = frames_v[0].nbytes
    vp = (0, 0, W, H) if full_clear else (10, 6, 160, 120)

    # reference: one frame at a time
    api = reference.api
    want = []
    for f, v in enumerate(frames_v):
        api.glInit(W, H)
        reference.lib.swglref_fill(0, C.c_float(0.0))
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, v, [(0, 4, 0), (1, 4, 16)])
        api.glClearColor(0.1, 0.2, 0.3, 1.0)
        api.glViewport(*vp)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(v))
        want.append(G.frame_color(api, W, H).copy())

    api = gpu_api
    api.glInit(W, H)
    api.swglFillFramebuffer(0, C.c_float(0.0))
    p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
    api.glUseProgram(p)
    sets = _two_buffer_sets(api, nbytes, [(0, 4, 0), (1, 4, 16)])
    api.glClearColor(0.1, 0.2, 0.3, 1.0)
    api.glViewport(*vp)
    w0 = api.swglGetOption(b"wt_draws")
    got, prev = [], 0

    def collect(ticket):
        ptr = api.swglFrameWait(ticket)
        assert ptr
        got.append(np.ctypeslib.as_array(ptr, shape=(H, W)).copy())

    for f, v in enumerate(frames_v):
        vao, vbo = sets[f & 1]
        api.glBindVertexArray(vao)
        api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo)
        api.swglBufferRespecify(G.GL_ARRAY_BUFFER, nbytes, _ptr(v))
        if not full_clear:
            api.swglFillFramebuffer(0, C.c_float(0.0))     # what glInit leaves outside the viewport in the reference run
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(v))
        t = api.swglFrameSubmit()
        assert t == f + 1
        if prev:
            collect(prev)
        prev = t
    collect(prev)
    assert api.swglGetLastError().decode() == ""
    assert api.swglGetOption(b"wt_draws") == w0      # pipelined frames travel by DMA, not write-through
    for f in range(len(frames_v)):
        assert np.array_equal(got[f], want[f]), f
    # a ticket older than the two mirrors is refused
    assert not api.swglFrameWait(1)
    assert b"ticket" in api.swglGetLastError()


def test_respecify_right_after_a_draw_keeps_the_queued_draw_intact(gpu_api, reference):
    """The overlapped upload must wait for the draws that still read the buffer (single-buffered client)."""
    a = np.ascontiguousarray(S.random_triangles(3000, W, H, seed=31, alpha=None, extent=0.6).vertices)
    b = np.ascontiguousarray(S.random_triangles(3000, W, H, seed=32, alpha=None, extent=0.6).vertices)

    def script(api, respec):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, a, [(0, 4, 0), (1, 4, 16)])
        api.glClear(3)
        for _ in range(4):
            api.glDrawArrays(G.GL_TRIANGLES, 0, len(a))
        respec(api)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(b))

    def gpu_respec(api):
        api.swglBufferRespecify(G.GL_ARRAY_BUFFER, b.nbytes, _ptr(b))

    def ref_respec(api):     # the reference ignores re-specification: a fresh buffer instead
        _vao(api, b, [(0, 4, 0), (1, 4, 16)])

    out = []
    for api, is_ref in ((gpu_api, False), (reference.api, True)):
        api.glInit(W, H)
        (reference.lib.swglref_fill if is_ref else api.swglFillFramebuffer)(0, C.c_float(0.0))
        api.glViewport(0, 0, W, H)
        api.glClearColor(0.0, 0.0, 0.0, 1.0)
        script(api, ref_respec if is_ref else gpu_respec)
        out.append(G.frame_color(api, W, H).copy())
    assert np.array_equal(out[0], out[1])


def test_read_pixels_rgba8_and_ppm(gpu_api, reference, tmp_path):
    scene = S.random_triangles(200, W, H, seed=5, alpha=None, extent=0.5)
    def script(api):
        st = G.setup_scene(api, scene, indexed=False, init=False)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)
    want_bytes = np.stack([(b[0] >> 24) & 255, (b[0] >> 16) & 255, (b[0] >> 8) & 255, b[0] & 255], axis=-1).astype(np.uint8)
    rgba = np.zeros((H, W, 4), np.uint8)
    assert gpu_api.swglReadPixelsRGBA8(rgba.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(rgba, want_bytes)
    path = tmp_path / "frame.ppm"
    assert gpu_api.swglWritePPM(str(path).encode()) == 0
    raw = path.read_bytes()
    header = f"P6\n{W} {H}\n255\n".encode()
    assert raw.startswith(header) and len(raw) == len(header) + W * H * 3
    assert np.array_equal(np.frombuffer(raw[len(header):], np.uint8).reshape(H, W, 3), want_bytes[:, :, :3])
    assert gpu_api.swglWritePPM(b"/nonexistent-dir/x.ppm") != 0
    assert b"swglWritePPM" in gpu_api.swglGetLastError()


@pytest.mark.parametrize("name", ["colour_near", "colour_viewport", "textured", "matrix"])
@pytest.mark.parametrize("lod", [0, 1])
def test_points_match_the_restatement(gpu_api, restatement, name, lod):
    """GL_POINTS against the C restatement (swglo_draw_points, pinned on the CPU against the compiled reference,
    tests/test_oracle.py), also with mip levels sampled: a point takes the level the last triangle left."""
    from test_oracle import _point_scenes
    from util import assert_bit_exact, gpu_render
    scene = _point_scenes()[name]
    if lod and scene.texture is None:
        pytest.skip("no texture")
    kw = dict(count=300, points=(300, len(scene.vertices) - 300), mipmaps=bool(lod))
    rc, rd, _ = restatement.render(scene, **kw)
    col, dep, stats, err = gpu_render(gpu_api, scene, indexed=False, options={"mip_lod": lod}, **kw)
    assert err == "", err
    assert_bit_exact(O.compare(col, dep, rc, rd), name)
