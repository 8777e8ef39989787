"""GPU: API-level edge cases, the same call sequence run against libswgl_b200.so and against the
compiled reference (oracle/_ref); frames (colour words and depth bits) must be identical.

Only behaviour that is DEFINED in the reference is exercised (no out-of-bounds buffer reads, no
reads of never-written variables, vec4 fragment output)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import gl as G, scenes as S

pytestmark = pytest.mark.gpu

W, H = 200, 152


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _program(api, vs, fs):
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
    return p, v, f


def _vao(api, verts, attribs, stride=None, bind_before_vao=False):
    verts = np.ascontiguousarray(verts, np.float32)
    vao, vbo = C.c_uint32(0), C.c_uint32(0)
    api.glGenBuffers(1, C.byref(vbo))
    if bind_before_vao:
        # no VAO bound: the named buffer itself receives the data (swgl.c:3123-3126) ...
        api.glBindVertexArray(0)
        api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
        api.glBufferData(G.GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), G.GL_STATIC_DRAW)
    api.glGenVertexArrays(1, C.byref(vao))
    api.glBindVertexArray(vao.value)
    api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)      # ... and is aliased into the VAO here (3116-3122)
    if not bind_before_vao:
        api.glBufferData(G.GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), G.GL_STATIC_DRAW)
    st = verts.shape[1] * 4 if stride is None else stride
    for loc, n, off in attribs:
        api.glVertexAttribPointer(loc, n, G.GL_FLOAT, G.GL_FALSE, st, C.c_void_p(off))
    return vao.value


def _both(gpu_api, reference, script):
    """Run script(api, is_ref) on both libraries; return ((col, dep), (col, dep))."""
    out = []
    for api, is_ref in ((gpu_api, False), (reference.api, True)):
        api.glInit(W, H)
        if is_ref:
            reference.lib.swglref_fill(0x0A0B0C0D, C.c_float(0.0))
        else:
            api.swglFillFramebuffer(0x0A0B0C0D, C.c_float(0.0))
        api.glViewport(0, 0, W, H)
        api.glClearColor(0.0, 0.0, 0.0, 1.0)
        script(api)
        col = G.frame_color(api, W, H)
        if is_ref:
            dep = np.ctypeslib.as_array(reference.lib.swglref_depth_ptr(), shape=(H, W)).copy()
        else:
            dep = np.ctypeslib.as_array(api.swglGetDepthPtr(), shape=(H, W)).copy()
            assert api.swglGetLastError().decode() == ""
        out.append((col, dep))
    return out


def _assert_same(a, b, min_covered=1):
    cmp = O.compare(a[0], a[1], b[0], b[1])
    assert cmp["color_mismatch"] == 0 and cmp["depth_mismatch"] == 0 and cmp["coverage_mismatch"] == 0, cmp
    assert int((b[1].view(np.uint32) != 0).sum()) >= min_covered


SCENE = S.random_triangles(120, W, H, seed=4242, alpha=None, extent=0.5)
VERTS = SCENE.vertices


def test_count_not_multiple_of_three_and_first_offset(gpu_api, reference):
    def script(api):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, VERTS, [(0, 4, 0), (1, 4, 16)])
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 6, 31)     # 11 triangles: the last one is a partial triple
        api.glDrawArrays(G.GL_TRIANGLES, 90, 3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, 0)      # nothing
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


def test_vec3_attribute_with_explicit_w_and_constant_colour(gpu_api, reference):
    vs = ("layout (location = 0) vec3 aPos;\nvoid main()\n{\n"
          "gl_Position = vec4(aPos.x, aPos.y, aPos.z, 1.0);\n}\n")
    fs = "out vec4 FragColor;\nvoid main()\n{\nFragColor = vec4(0.25, 0.5, 0.75, 0.5);\n}\n"
    v = np.ascontiguousarray(VERTS[:, :3] / VERTS[:, 3:4])

    def script(api):
        p, _, _ = _program(api, vs, fs)
        api.glUseProgram(p)
        _vao(api, v, [(0, 3, 0)])
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(v))
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


def test_duplicate_location_last_pointer_wins_and_stride_zero_colour(gpu_api, reference):
    def script(api):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        vao = _vao(api, VERTS, [(0, 4, 0), (1, 4, 0), (1, 4, 16)])   # location 1 fed twice
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, 150)
        # a second VAO on the same data whose colour attribute has stride 0 (every vertex: first colour)
        _vao(api, VERTS, [(0, 4, 0)])
        api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 0, C.c_void_p(16))
        api.glDrawArrays(G.GL_TRIANGLES, 150, 60)
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


def test_buffer_specified_before_vao_and_respecification_ignored(gpu_api, reference):
    other = np.zeros_like(VERTS)

    def script(api):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, VERTS, [(0, 4, 0), (1, 4, 16)], bind_before_vao=True)
        api.glBufferData(G.GL_ARRAY_BUFFER, other.nbytes, _ptr(other), G.GL_STATIC_DRAW)  # ignored (swgl.c:3140)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(VERTS))
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


def test_uniform_updates_between_draws_and_shared_shader_objects(gpu_api, reference):
    vs = ("layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\nuniform vec4 shift;\nout vec4 vCol;\n"
          "void main()\n{\ngl_Position = aPos + shift;\nvCol = aCol;\n}\n")
    fs_a = "in vec4 vCol;\nuniform vec4 tint;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol * tint;\n}\n"
    fs_b = "in vec4 vCol;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = vCol.zyxw;\n}\n"

    def script(api):
        pa, v, fa = _program(api, vs, fs_a)
        fb = api.glCreateShader(G.GL_FRAGMENT_SHADER)
        api.glShaderSource(fb, fs_b.encode())
        api.glCompileShader(fb)
        pb = api.glCreateProgram()
        api.glAttachShader(pb, v)          # the SAME vertex shader object: uniform storage is shared
        api.glAttachShader(pb, fb)
        api.glLinkProgram(pb)
        _vao(api, VERTS, [(0, 4, 0), (1, 4, 16)])
        api.glClear(3)
        api.glUseProgram(pa)
        api.glUniform4f(api.glGetUniformLocation(pa, b"shift"), 0.1, -0.2, 0.0, 0.0)
        api.glUniform4f(api.glGetUniformLocation(pa, b"tint"), 1.0, 0.5, 0.25, 0.5)
        api.glDrawArrays(G.GL_TRIANGLES, 0, 120)
        api.glUniform4f(api.glGetUniformLocation(pa, b"tint"), 0.2, 0.9, 0.4, 1.0)
        api.glDrawArrays(G.GL_TRIANGLES, 120, 120)
        api.glUseProgram(pb)               # sees shift = (0.1, -0.2, 0, 0) set through program A
        api.glDrawArrays(G.GL_TRIANGLES, 240, 120)
        api.glUniform1f(api.glGetUniformLocation(pb, b"shift"), 3.0)   # wrong type: ignored
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


def test_no_program_or_no_vertex_array_draws_nothing(gpu_api, reference):
    def script(api):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        vao = _vao(api, VERTS, [(0, 4, 0), (1, 4, 16)])
        api.glClear(3)
        api.glUseProgram(0)
        api.glDrawArrays(G.GL_TRIANGLES, 0, 90)    # no program: returns (swgl.c:3478)
        api.glUseProgram(p)
        api.glBindVertexArray(0)
        api.glDrawArrays(G.GL_TRIANGLES, 0, 90)    # no VAO: returns (swgl.c:3477)
        api.glBindVertexArray(vao)
        api.glDrawArrays(G.GL_TRIANGLES, 90, 30)
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 10)


def test_partial_clears_and_clear_colour_clamp(gpu_api, reference):
    def script(api):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, VERTS, [(0, 4, 0), (1, 4, 16)])
        api.glClearColor(1.5, -0.5, 0.3, 0.999)           # clamped (swgl.c:3177-3180)
        api.glClear(G.GL_COLOR_BUFFER_BIT)                 # depth keeps the fill value
        api.glDrawArrays(G.GL_TRIANGLES, 0, 150)
        api.glViewport(20, 10, 90, 70)
        api.glClearColor(0.1, 0.2, 0.3, 0.4)
        api.glClear(G.GL_DEPTH_BUFFER_BIT)                 # depth only, inside the small viewport
        api.glDrawArrays(G.GL_TRIANGLES, 150, 150)
        api.glClear(G.GL_COLOR_BUFFER_BIT)
        api.glClear(G.GL_DEPTH_BUFFER_BIT)                 # two clears merge
        api.glViewport(0, 0, W, H)
        api.glDrawArrays(G.GL_TRIANGLES, 300, 60)
        api.glViewport(7, 3, 121, 97)                      # odd sizes: VW/2 is an integer division
        api.glClear(3)
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


def test_two_textures_on_other_units(gpu_api, reference):
    fs = ("in vec4 vCol;\nuniform sampler2D texA;\nuniform sampler2D texB;\nout vec4 FragColor;\nvoid main()\n{\n"
          "vec4 a = texture(texA,vCol.xy);\nvec4 b = texture(texB,vCol.zy);\nFragColor = a * b + vCol * vec4(0.1, 0.1, 0.1, 0.0);\n}\n")
    ta = S.checker_texture(32)
    tb = S.lcg_texture(16, seed=5)[:, :, :3].copy()        # RGB8: alpha reads 0 (swgl.c:2561)

    def script(api):
        p, _, _ = _program(api, S.VS_PASSTHROUGH, fs)
        api.glUseProgram(p)
        _vao(api, VERTS, [(0, 4, 0), (1, 4, 16)])
        t1, t2 = C.c_uint32(0), C.c_uint32(0)
        api.glGenTextures(1, C.byref(t1))
        api.glGenTextures(1, C.byref(t2))
        api.glActiveTexture(G.GL_TEXTURE3)
        api.glBindTexture(G.GL_TEXTURE_2D, t1.value)
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 32, 32, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(ta))
        api.glActiveTexture(G.GL_TEXTURE5)
        api.glBindTexture(G.GL_TEXTURE_2D, t2.value)
        api.glTexParameteri(G.GL_TEXTURE_2D, G.GL_TEXTURE_WRAP_S, G.GL_CLAMP)
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGB, 16, 16, 0, G.GL_RGB, G.GL_UNSIGNED_BYTE, _ptr(tb))
        api.glUniform1i(api.glGetUniformLocation(p, b"texA"), 3)
        api.glUniform1i(api.glGetUniformLocation(p, b"texB"), 5)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(VERTS))
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


@pytest.mark.parametrize("path", [1, 2, 3], ids=["pixel_owner", "fragment_parallel", "warp_tile"])
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_extreme_coordinates_and_w(gpu_api, reference, seed, path):
    """Huge / tiny / negative w, vertices far off screen, zero-area triangles: NaN and INT_MIN paths
    (x86 cvttss2si and NaN bit patterns)."""
    rng = np.random.default_rng(seed)
    n = 400
    v = np.empty((n * 3, 8), np.float32)
    v[:, 0:2] = rng.normal(0, 1.0, (n * 3, 2)) * rng.choice([0.3, 1.0, 30.0, 1e4], (n * 3, 1))
    v[:, 2] = rng.uniform(-1.5, 1.5, n * 3)
    v[:, 3] = rng.choice([1.0, 0.5, 2.0, 1e-3, -1.0, 1e3, 0.25], n * 3) * rng.uniform(0.5, 1.5, n * 3)
    v[:, 4:8] = rng.uniform(-0.2, 1.2, (n * 3, 4))
    v[30:33] = v[30]                      # zero-area triangle
    v[60:63, 0:2] = v[60, 0:2]            # collapsed in x, y only
    v[90:93, 3] = 0.0                     # w = 0: division by zero in the viewport transform

    def script(api):
        if api is gpu_api:
            api.swglSetOption(b"raster_path", path)
        p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
        api.glUseProgram(p)
        _vao(api, v, [(0, 4, 0), (1, 4, 16)])
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(v))
    a, b = _both(gpu_api, reference, script)
    _assert_same(a, b, 100)


def test_texture_respecified_right_after_a_draw_that_overflowed(gpu_api, reference):
    """A draw whose tile lists overflowed is dropped on the device and re-issued later from the
    pointers it was launched with: glTexImage2D on the texture it samples (free + upload) must
    resolve it first (the reference has drawn with the old texels by then)."""
    ta = S.lcg_texture(32, seed=9)
    tb = 255 - ta
    fs = "in vec4 vCol;\nuniform sampler2D tex;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = texture(tex,vCol.xy);\n}\n"

    def script(api):
        if api is gpu_api:
            api.swglSetOption(b"bin_cap", 4)         # every list overflows on the first attempt
        p, _, _ = _program(api, S.VS_PASSTHROUGH, fs)
        api.glUseProgram(p)
        _vao(api, VERTS, [(0, 4, 0), (1, 4, 16)])
        t = C.c_uint32(0)
        api.glGenTextures(1, C.byref(t))
        api.glActiveTexture(G.GL_TEXTURE0)
        api.glBindTexture(G.GL_TEXTURE_2D, t.value)
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 32, 32, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(ta))
        api.glUniform1i(api.glGetUniformLocation(p, b"tex"), 0)
        api.glClear(3)
        api.glDrawArrays(G.GL_TRIANGLES, 0, len(VERTS))
        # same size: the allocator is likely to hand the freed block straight back
        api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, 32, 32, 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(tb))
    a, b = _both(gpu_api, reference, script)
    assert gpu_api.swglGetOption(b"bin_cap") > 4
    _assert_same(a, b, 100)


def test_respecify_that_moves_the_storage_reaches_every_vertex_array_sharing_it(gpu_api, restatement):
    """glBindBuffer snapshots the named buffer into the vertex array's own struct (swgl.c:3116-3122); two vertex
    arrays that bound the same name share its storage.  When swglBufferRespecify through one of them has to
    grow (move) the storage, the other one must follow instead of reading freed memory."""
    small, big = S.random_triangles(60, W, H, seed=11), S.random_triangles(200, W, H, seed=12)
    api = gpu_api
    api.glInit(W, H)
    api.swglFillFramebuffer(0, C.c_float(0.0))
    p, _, _ = _program(api, S.VS_PASSTHROUGH, S.FS_COLOR)
    api.glUseProgram(p)
    api.glViewport(0, 0, W, H)
    api.glClearColor(0.0, 0.0, 0.0, 1.0)
    v0 = np.ascontiguousarray(small.vertices, np.float32)
    vbo = C.c_uint32(0)
    api.glGenBuffers(1, C.byref(vbo))
    api.glBindVertexArray(0)
    api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)             # no vertex array bound: the name owns the data
    api.glBufferData(G.GL_ARRAY_BUFFER, v0.nbytes, _ptr(v0), G.GL_STATIC_DRAW)
    vaos = []
    for _ in range(2):
        vao = C.c_uint32(0)
        api.glGenVertexArrays(1, C.byref(vao))
        api.glBindVertexArray(vao.value)
        api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)         # snapshot (alias) of the named buffer
        api.glVertexAttribPointer(0, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(0))
        api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(16))
        vaos.append(vao.value)
    v1 = np.ascontiguousarray(big.vertices, np.float32)
    api.glBindVertexArray(vaos[0])
    api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
    api.swglBufferRespecify(G.GL_ARRAY_BUFFER, v1.nbytes, _ptr(v1))     # larger: the storage moves
    api.glBindVertexArray(vaos[1])                             # the other array still knows the old size only
    api.glClear(3)
    api.glDrawArrays(G.GL_TRIANGLES, 0, len(v0))
    col = G.frame_color(api, W, H)
    assert api.swglGetLastError().decode() == ""
    want = restatement.render(big, count=len(v0))[0]
    assert np.array_equal(col, want)


@pytest.mark.parametrize("viewport", [(0, -40, W, H), (0, 30, W, H), (-10, -25, W + 20, H + 60), (13, -7, 150, 300),
                                      (89, -24, 101, 10), (0, H + 8, W, 50)],
                         ids=["below", "above", "both", "narrow_tall", "entirely_below_row_limit_wraps", "entirely_above"])
@pytest.mark.parametrize("kind", ["colour_alpha_nearclip", "textured_mesh", "generic_shader"])
def test_viewport_that_leaves_the_framebuffer_rows_folds_onto_the_last_row(gpu_api, reference, viewport, kind):
    """swgl.c:3386: raster rows whose storage row would fall outside the framebuffer all land on row Height-1,
    in (primitive, y) order.  The library renders the rows inside through the ordinary kernels (into a virtual
    framebuffer) and row Height-1 with k_fold_row; the frame must be the reference's, bit for bit."""
    if kind == "colour_alpha_nearclip":
        scene = S.random_triangles(400, W, H, seed=21, near_cross=True, alpha=None, centre_range=1.2)
    elif kind == "textured_mesh":
        scene = S.grid_mesh(20, W, H, textured=True)
    else:
        scene = S.random_triangles(300, W, H, seed=22, alpha=None)
        scene.fs = ("in vec4 vCol;\nuniform vec4 tint;\nout vec4 FragColor;\nvoid main()\n{\n"
                    "FragColor = vCol * tint + vec4(0.1, 0.0, 0.2, 0.0);\n}\n")
        scene.uniforms = {"tint": ("4f", [0.9, 0.7, 0.8, 0.6])}

    def script(api):
        st = G.setup_scene(api, scene, indexed=False, init=False)
        api.glViewport(0, 0, W, H)
        api.glClear(3)
        api.glViewport(*viewport)
        api.glClear(G.GL_DEPTH_BUFFER_BIT)           # partial: viewport ∩ framebuffer, rows not flipped
        api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
        api.glDrawArrays(G.GL_TRIANGLES, 3, st["n_draw"] - 3)     # a second draw blends over the folded row again
    a, b = _both(gpu_api, reference, script)
    assert gpu_api.swglGetOption(b"draws_folded") == 2 and gpu_api.swglGetOption(b"draws_refused") == 0
    # (a viewport that ends below row 0: VY + VH wraps in the reference's unsigned arithmetic, swgl.c:3344 / 3356 -- the row
    # limit is gone, every row of every triangle lands on row Height-1, and the depth clear covers every framebuffer row)
    outside_all = viewport[1] + viewport[3] <= 0 or viewport[1] >= H
    _assert_same(a, b, 0 if outside_all else 100)
    # the last row really collects fragments of several raster rows (the mesh's triangles are less than a row tall in a
    # 10-row viewport: the reference draws none of them)
    if not (outside_all and kind == "textured_mesh"):
        assert int((b[1][H - 1].view(np.uint32) != 0).sum()) > 0


@pytest.mark.parametrize("viewport", [(0, -40, W, H), (0, 30, W, H), (-10, -25, W + 20, H + 60), (13, -7, 150, 300),
                                      (89, -24, 101, 10), (0, H + 8, W, 50)],
                         ids=["below", "above", "both", "narrow_tall", "entirely_below_row_limit_wraps", "entirely_above"])
def test_folded_draw_counts_fragments_like_the_restatement(gpu_api, restatement, viewport):
    """The C restatement folds rows like the reference (checked against it on the CPU: tests/test_oracle.py, test_restatement_folds_rows_like_the_compiled_reference) and
    counts Barycentric calls and depth passes: a folded draw must report the same numbers -- every fragment tested once
    (by the tile kernels for the rows they walk, by k_fold_row beyond) and shaded on the real row it lands on."""
    from util import gpu_render
    scene = S.random_triangles(400, W, H, seed=21, near_cross=True, alpha=None, centre_range=1.2)
    scene.viewport = viewport
    rc, rd, rstats = restatement.render(scene)
    col, dep, stats, err = gpu_render(gpu_api, scene)
    assert err == "" and gpu_api.swglGetOption(b"draws_folded") == 1
    cmp = O.compare(col, dep, rc, rd)
    assert cmp["color_mismatch"] == 0 and cmp["depth_mismatch"] == 0 and cmp["coverage_mismatch"] == 0, cmp
    assert (stats["tested"], stats["shaded"]) == (rstats["tested"], rstats["shaded"])
