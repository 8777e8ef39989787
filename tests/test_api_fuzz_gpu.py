"""GPU: random SEQUENCES of the reference's own API calls -- programs, vertex arrays, textures, uniforms, viewports
(also ones that leave the framebuffer), partial clears, draws with arbitrary first / count, frames read back in the
middle -- issued identically to libswgl_b200.so and to the compiled reference (oracle/_ref).  Every frame read on the
way and the final colour words and depth bits must be identical.

Only behaviour that is DEFINED in the reference is generated (no reads past a buffer, well-typed shaders, vec4
fragment output).  tools/api_fuzz.py runs many more seeds of the same generator."""
import ctypes as C

import numpy as np
import pytest

import swgl_b200
from oracle import pyoracle as O
from swgl_b200 import gl as G, scenes as S

pytestmark = pytest.mark.gpu

SIZES = [(200, 152), (200, 152), (97, 75), (333, 211), (256, 64), (130, 300)]

VS_COLOUR = S.VS_PASSTHROUGH
FS_COLOUR = S.FS_COLOR
VS_MATRIX = ("layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\nuniform mat4 uM;\nout vec4 vCol;\n"
             "void main()\n{\ngl_Position = uM * aPos;\nvCol = aCol;\n}\n")
VS_UV = ("layout (location = 0) vec4 aPos;\nlayout (location = 1) vec4 aCol;\nout vec2 vUV;\n"
         "void main()\n{\ngl_Position = aPos;\nvUV = aCol.xy;\n}\n")
FS_TEX = "in vec2 vUV;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\nFragColor = texture(uTex, vUV);\n}\n"
FS_TINT = ("in vec4 vCol;\nuniform vec4 tint;\nuniform sampler2D uTex;\nout vec4 FragColor;\nvoid main()\n{\n"
           "vec4 t = texture(uTex,vCol.zy) * tint;\nFragColor = vec4(t.x, vCol.y, t.z, vCol.w);\n}\n")
PROGRAMS = [(VS_COLOUR, FS_COLOUR), (VS_MATRIX, FS_COLOUR), (VS_UV, FS_TEX), (VS_COLOUR, FS_TINT), (VS_MATRIX, FS_TINT)]
# ... and a pool of random well-typed programs (tests/shader_fuzz_gen.py): varyings of every width, uniforms u1 .. u4 that
# the sequences change under way; compiled at run time once per process, interpreted on the fold row
from shader_fuzz_gen import make as _random_program          # noqa: E402
RANDOM_PROGRAMS = [_random_program(900 + k) for k in range(6)]
PROGRAMS += [(vs, fs) for vs, fs, _ in RANDOM_PROGRAMS]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def make_ops(seed, inside=False, lod=False):
    """The call sequence of a seed as a list of tuples (data only: the same list drives both libraries).
    inside: viewports stay inside the framebuffer rows (sort-first ranks of separate processes cannot fold);
    lod: for runs that SAMPLE mip chains ("mip_lod"): one chain exists from the start, and a texture object that has
    levels never gets an image with MORE floats per texel than they were built with (the reference never clears a
    texture's levels and addresses them with the current image's floats per texel, swgl.c:2094-2118, 2559: with more of
    them it reads past the level's allocation).  Second chains appended behind the first and levels of images that
    have since been replaced are part of the game."""
    rng = np.random.default_rng(31000 + seed)
    W, H = SIZES[int(rng.integers(len(SIZES)))] if seed >= 12 else SIZES[0]      # odd sizes: partial tiles, unaligned rows

    def new_array():
        sc = S.random_triangles(int(rng.integers(40, 260)), W, H, seed=int(rng.integers(1, 1 << 30)),
                                extent=float(rng.choice([0.08, 0.25, 0.6, 1.3])),
                                alpha=None if rng.random() < 0.5 else float(rng.choice([1.0, 0.5, 0.15])),
                                near_cross=bool(rng.random() < 0.4), centre_range=float(rng.choice([0.8, 1.0, 1.3])))
        if rng.random() < 0.3:
            # an indexed mesh: the library draws it with glDrawElements, the reference (which has no indexed draw) draws
            # the de-indexed stream -- "glDrawArrays over the de-indexed vertex stream" is the definition of the extension
            sc = S.grid_mesh(int(rng.integers(3, 28)), W, H, seed=int(rng.integers(1, 1 << 30)), alpha=float(rng.choice([1.0, 0.5])),
                             layers=int(rng.integers(1, 3)))
            return (np.ascontiguousarray(sc.vertices, np.float32), np.ascontiguousarray(sc.indices, np.uint32))
        return (np.ascontiguousarray(sc.vertices, np.float32), None)

    arrays = [new_array() for _ in range(int(rng.integers(2, 4)))]
    textures = [S.checker_texture(int(rng.choice([8, 32]))), S.lcg_texture(int(rng.choice([16, 64])), seed=int(rng.integers(1, 99)))]
    ops = [("setup", list(arrays), textures, W, H)]
    prog, vao = 0, 0
    bound, level_fpp = 1, {}              # the texture object the texture calls act on: the one bound last (swgl.c:2062)
    cur_fpp = {0: 4, 1: 4}
    if lod:
        ops.append(("mipmap", 1))
        level_fpp[bound] = 4
    ops += [("use", 0), ("vao", 0), ("clear", 3)]
    for _ in range(int(rng.integers(10, 26))):
        r = rng.random()
        if rng.random() < 0.06:
            # new contents for a vertex array's buffers (the streaming client of bench.py's end-to-end step): the library
            # re-specifies in place (swglBufferRespecify, the reference ignores re-specification, swgl.c:3140), the
            # reference gets a fresh vertex array -- same geometry from here on either way
            vao = int(rng.integers(len(arrays)))
            new = new_array()
            if (new[1] is None) != (arrays[vao][1] is None):          # keep the array's kind (indexed or not)
                new = (new[0], None) if arrays[vao][1] is None else (arrays[vao][0], arrays[vao][1][: 3 * int(rng.integers(1, len(arrays[vao][1]) // 3 + 1))].copy())
            arrays[vao] = new
            ops.append(("respecify", vao, new))
        elif r < 0.04 and arrays[vao][1] is None:
            n = len(arrays[vao][0])
            first = int(rng.integers(0, n - 1))
            ops.append(("points", first, int(rng.integers(0, min(n - first, 60) + 1))))
        elif r < 0.07:
            unit = int(rng.integers(0, 2))
            tex = S.lcg_texture(int(rng.choice([4, 16, 32])), seed=int(rng.integers(1, 99)))
            tex = tex if rng.random() < 0.7 else np.ascontiguousarray(tex[:, :, :3])
            if rng.random() < 0.3:
                bound = int(rng.integers(0, 2))
                ops.append(("bindtex", unit, bound))      # texture object `bound` onto `unit`; it becomes the one texture calls act on
            if not (lod and tex.shape[2] > level_fpp.get(bound, 4)):
                ops.append(("teximage", unit, tex))
                cur_fpp[bound] = tex.shape[2]
            if rng.random() < 0.4:
                ops.append(("mipmap", unit))
                level_fpp[bound] = min(level_fpp.get(bound, 4), cur_fpp[bound])
        elif r < 0.38:
            n = len(arrays[vao][0]) if arrays[vao][1] is None else len(arrays[vao][1])
            first = int(rng.integers(0, max(n - 2, 1)))
            if rng.random() < 0.8:
                first -= first % 3
            count = int(rng.integers(0, n - first + 1))
            if first + 3 * ((count + 2) // 3) > n:        # a partial last triple is read whole: keep it inside the buffer
                count = 3 * ((n - first) // 3)
            ops.append(("draw", first, count))
        elif r < 0.50:
            if rng.random() < 0.5:
                ops.append(("viewport", 0, 0, W, H))
            else:
                x = int(rng.integers(-24, W // 2)) if rng.random() < 0.3 else int(rng.integers(0, W // 2))
                y = int(rng.integers(-40, H // 2)) if rng.random() < 0.3 and not inside else int(rng.integers(0, H // 2))
                w = int(rng.integers(8, W - max(x, 0) + (24 if rng.random() < 0.3 else 0) + 1))
                h = int(rng.integers(8, H - max(y, 0) + (48 if rng.random() < 0.3 and not inside else 0) + 1))
                ops.append(("viewport", x, y, w, h))
        elif r < 0.60:
            ops.append(("clearcolor",) + tuple(float(v) for v in rng.uniform(-0.2, 1.2, 4)))
            ops.append(("clear", int(rng.integers(1, 4))))
        elif r < 0.70:
            prog = int(rng.integers(len(PROGRAMS)))
            ops.append(("use", prog))
        elif r < 0.78:
            vao = int(rng.integers(len(arrays)))
            ops.append(("vao", vao))
        elif r < 0.86:
            m = np.eye(4, dtype=np.float32) + rng.uniform(-0.15, 0.15, (4, 4)).astype(np.float32)
            ops.append(("matrix", m))
        elif r < 0.91:
            if rng.random() < 0.5:
                ops.append(("tint",) + tuple(float(v) for v in rng.uniform(0.1, 1.3, 4)))
            else:
                n = int(rng.integers(1, 5))                   # u1 (float) .. u4 (vec4) of the random programs
                ops.append(("uni", n) + tuple(float(v) for v in rng.uniform(0.1, 1.5, n)))
        elif r < 0.95:
            ops.append(("sampler", int(rng.integers(0, 2))))
        elif r < 0.97:
            ops.append(("wrap", int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 2))))
        else:
            ops.append(("read",))
    return ops


PERTURB = [("host_mirror", (0, 1, 2)), ("fuse_clear", (0, 1)), ("tile_rows", (0, 8, 4, 2)), ("raster_path", (0, 2, 3)),
           ("lean_prims", (0, 1)), ("setup_big", (0, 1)), ("jit", (0, 1)), ("overflow_pool", (0, 1)), ("bin_cap", (8, 256)),
           ("count_fragments", (0, 1)), ("finish",), ("stats",), ("submit_wait",), ("rgba8",)]


def run_ops(api, ops, fill, depth_of, perturb=None, devices=1, ours=False, points=True, options=None, stripe=None):
    """Issue `ops`; -> (frames read on the way + the final one, final depth).
    perturb: seed of library-only calls slipped in between (options that must not change a bit of the result, waits,
    statistics, the pipelined and the byte-swizzled read-back); devices: swglSetDeviceCount before glInit;
    ours: the library has glDrawElements and swglBufferRespecify (the reference draws an indexed array as its
    de-indexed stream and gets a fresh vertex array where the library re-specifies one)."""
    W, H = ops[0][3], ops[0][4]
    prng = np.random.default_rng(perturb) if perturb is not None else None
    # (folding is a feature of the default rasteriser: the CTA cross-check kernel refuses such draws, by design)
    leaves_rows = any(o[0] == "viewport" and (o[2] < 0 or o[2] + o[4] > H) for o in ops)
    if devices > 1:
        api.swglSetDeviceCount(devices)
    api.glInit(W, H)
    if devices > 1:
        api.swglSetDeviceCount(1)          # the setting is consumed by glInit: later tests get one device again
    for name, value in (options or {}).items():      # device options start from their defaults at every glInit
        api.swglSetOption(name.encode(), value)
    if stripe:
        api.swglSetStripe(*stripe)                   # (rank, ranks, band height in 32-row units): sort-first share of one rank
    fill(0x0A0B0C0D, 0.0)
    api.glViewport(0, 0, W, H)
    api.glClearColor(0.0, 0.0, 0.0, 1.0)
    frames, progs, vaos, cur, has_ebo, cur_vao, names, texs = [], [], [], 0, [], 0, [], []

    def make_array(j, verts, idx):
        """vertex array j from scratch: named buffers that own the data (specified with no vertex array bound, swgl.c:3123-3126),
        then bound into a new vertex array (which snapshots their fields, 3116-3122)"""
        if idx is not None and not ours:
            verts = np.ascontiguousarray(verts[idx])
        vao, vbo, ebo = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        api.glBindVertexArray(0)
        api.glGenBuffers(1, C.byref(vbo)); api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
        api.glBufferData(G.GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts), G.GL_STATIC_DRAW)
        if idx is not None and ours:
            api.glGenBuffers(1, C.byref(ebo)); api.glBindBuffer(G.GL_ELEMENT_ARRAY_BUFFER, ebo.value)
            api.glBufferData(G.GL_ELEMENT_ARRAY_BUFFER, idx.nbytes, _ptr(idx), G.GL_STATIC_DRAW)
        api.glGenVertexArrays(1, C.byref(vao)); api.glBindVertexArray(vao.value)
        api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
        api.glVertexAttribPointer(0, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(0))
        api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(16))
        if ebo.value:
            api.glBindBuffer(G.GL_ELEMENT_ARRAY_BUFFER, ebo.value)
        vaos[j], names[j], has_ebo[j] = vao.value, (vbo.value, ebo.value), bool(ebo.value)

    for op in ops:
        k = op[0]
        if prng is not None and k != "setup" and prng.random() < 0.35:
            pk = PERTURB[int(prng.integers(len(PERTURB)))]
            if pk[0] == "finish":
                api.swglFinish()
            elif pk[0] == "stats":
                st = swgl_b200.swglStats()
                api.swglGetStats(C.byref(st))
            elif pk[0] == "submit_wait" and devices == 1:
                p = api.swglFrameWait(api.swglFrameSubmit())
                assert bool(p)
            elif pk[0] == "rgba8":
                buf = np.empty((H, W, 4), np.uint8)
                api.swglReadPixelsRGBA8(_ptr(buf))
            elif len(pk) == 2:
                v = int(pk[1][int(prng.integers(len(pk[1])))])
                if not (pk[0] == "raster_path" and v == 2 and leaves_rows):
                    api.swglSetOption(pk[0].encode(), v)
        if k == "setup":
            for vs, fs in PROGRAMS:
                v = api.glCreateShader(G.GL_VERTEX_SHADER); api.glShaderSource(v, vs.encode()); api.glCompileShader(v)
                f = api.glCreateShader(G.GL_FRAGMENT_SHADER); api.glShaderSource(f, fs.encode()); api.glCompileShader(f)
                p = api.glCreateProgram(); api.glAttachShader(p, v); api.glAttachShader(p, f); api.glLinkProgram(p)
                progs.append(p)
                # every uniform gets a value before the first draw (the reference's storage starts uninitialised)
                api.glUseProgram(p)
                ident = np.eye(4, dtype=np.float32)
                for name, setter in ((b"uM", lambda l: api.glUniformMatrix4fv(l, 1, G.GL_FALSE, ident.ctypes.data_as(C.POINTER(C.c_float)))),
                                     (b"tint", lambda l: api.glUniform4f(l, 1.0, 1.0, 1.0, 1.0)), (b"uTex", lambda l: api.glUniform1i(l, 0))):
                    loc = api.glGetUniformLocation(p, name)
                    if loc >= 0:
                        setter(loc)
                k = len(progs) - 1 - (len(PROGRAMS) - len(RANDOM_PROGRAMS))
                if k >= 0:
                    for name, (kind, vals) in RANDOM_PROGRAMS[k][2].items():
                        getattr(api, "glUniform" + kind)(api.glGetUniformLocation(p, name.encode()), *[float(x) for x in vals])
            for verts, idx in op[1]:
                vaos.append(None); names.append(None); has_ebo.append(False)
                make_array(len(vaos) - 1, verts, idx)
            for unit, tex in enumerate(op[2]):
                t = C.c_uint32(0)
                api.glGenTextures(1, C.byref(t))
                texs.append(t.value)
                api.glActiveTexture(G.GL_TEXTURE0 + unit)
                api.glBindTexture(G.GL_TEXTURE_2D, t.value)
                tt = np.ascontiguousarray(tex, np.uint8)
                api.glTexImage2D(G.GL_TEXTURE_2D, 0, G.GL_RGBA, tt.shape[1], tt.shape[0], 0, G.GL_RGBA, G.GL_UNSIGNED_BYTE, _ptr(tt))
            # the reference allocates a program's fragment `in` variables when it first shades a TRIANGLE fragment, and its
            # GL_POINTS path copies into them unconditionally (swgl.c:3548-3551): one certain fragment per program first
            warm = np.array([[-0.5, -0.5, -1.0, 1.0, 0.2, 0.4, 0.6, 1.0], [0.5, -0.5, -1.0, 1.0, 0.2, 0.4, 0.6, 1.0],
                             [0.0, 0.5, -1.0, 1.0, 0.2, 0.4, 0.6, 1.0]], np.float32)
            vao, vbo = C.c_uint32(0), C.c_uint32(0)
            api.glGenVertexArrays(1, C.byref(vao)); api.glBindVertexArray(vao.value)
            api.glGenBuffers(1, C.byref(vbo)); api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo.value)
            api.glBufferData(G.GL_ARRAY_BUFFER, warm.nbytes, _ptr(warm), G.GL_STATIC_DRAW)
            api.glVertexAttribPointer(0, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(0))
            api.glVertexAttribPointer(1, 4, G.GL_FLOAT, G.GL_FALSE, 32, C.c_void_p(16))
            for p in progs:
                api.glUseProgram(p)
                api.glDrawArrays(G.GL_TRIANGLES, 0, 3)
        elif k == "use":
            cur = progs[op[1]]
            api.glUseProgram(cur)
        elif k == "vao":
            cur_vao = op[1]
            api.glBindVertexArray(vaos[cur_vao])
        elif k == "respecify":
            cur_vao = op[1]
            verts, idx = op[2]
            if ours:
                api.glBindVertexArray(vaos[cur_vao])
                api.glBindBuffer(G.GL_ARRAY_BUFFER, names[cur_vao][0])
                api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.nbytes, _ptr(verts))
                if has_ebo[cur_vao]:
                    api.glBindBuffer(G.GL_ELEMENT_ARRAY_BUFFER, names[cur_vao][1])
                    api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, idx.nbytes, _ptr(idx))
            else:
                make_array(cur_vao, verts, idx)
        elif k == "draw":
            if has_ebo[cur_vao]:
                api.glDrawElements(G.GL_TRIANGLES, op[2], G.GL_UNSIGNED_INT, C.c_void_p(4 * op[1]))
            else:
                api.glDrawArrays(G.GL_TRIANGLES, op[1], op[2])
        elif k == "points" and points:
            api.glDrawArrays(G.GL_POINTS, op[1], op[2])
        elif k == "teximage":
            api.glActiveTexture(G.GL_TEXTURE0 + op[1])
            tt = np.ascontiguousarray(op[2], np.uint8)
            fmt = G.GL_RGBA if tt.shape[2] == 4 else G.GL_RGB
            api.glTexImage2D(G.GL_TEXTURE_2D, 0, fmt, tt.shape[1], tt.shape[0], 0, fmt, G.GL_UNSIGNED_BYTE, _ptr(tt))
        elif k == "mipmap":
            api.glActiveTexture(G.GL_TEXTURE0 + op[1])
            api.glGenerateMipmap(G.GL_TEXTURE_2D)
        elif k == "bindtex":
            api.glActiveTexture(G.GL_TEXTURE0 + op[1])
            api.glBindTexture(G.GL_TEXTURE_2D, texs[op[2]])
        elif k == "viewport":
            api.glViewport(*op[1:])
        elif k == "clearcolor":
            api.glClearColor(*op[1:])
        elif k == "clear":
            api.glClear(op[1])
        elif k == "matrix":
            loc = api.glGetUniformLocation(cur, b"uM")
            if loc >= 0:
                api.glUniformMatrix4fv(loc, 1, G.GL_FALSE, op[1].ctypes.data_as(C.POINTER(C.c_float)))
        elif k == "tint":
            loc = api.glGetUniformLocation(cur, b"tint")
            if loc >= 0:
                api.glUniform4f(loc, *op[1:])
        elif k == "uni":
            loc = api.glGetUniformLocation(cur, f"u{op[1]}".encode())
            if loc >= 0:
                getattr(api, f"glUniform{op[1]}f")(loc, *op[2:])
        elif k == "sampler":
            loc = api.glGetUniformLocation(cur, b"uTex")
            if loc >= 0:
                api.glUniform1i(loc, op[1])
        elif k == "wrap":
            api.glActiveTexture(G.GL_TEXTURE0 + op[1])
            api.glTexParameteri(G.GL_TEXTURE_2D, G.GL_TEXTURE_WRAP_S if op[2] else G.GL_TEXTURE_WRAP_T, G.GL_CLAMP if op[3] else G.GL_REPEAT)
        elif k == "read":
            frames.append(G.frame_color(api, W, H))
    frames.append(G.frame_color(api, W, H))
    return frames, depth_of()


def compare_seed(gpu_api, reference, seed, perturb=False, devices=1, lod=False):
    """-> '' if the two libraries agree on every frame of the sequence, else a description.
    lod: the library samples mip chains with the per-triangle level ("mip_lod" = 1) and `reference` is the build with
    the defined rsqrt (oracle/ref_shim.c); GL_POINTS draws sample with the level the last triangle left (a global in
    the reference; k_last_level here)."""
    ops = make_ops(seed, lod=lod)
    return compare_ops(gpu_api, reference, ops, f"seed {seed}", perturb=5000 + seed if perturb else None, devices=devices, lod=lod)


def compare_ops(gpu_api, reference, ops, what, perturb=None, devices=1, lod=False):
    W, H = ops[0][3], ops[0][4]
    try:
        gf, gd = run_ops(gpu_api, ops, lambda w, d: gpu_api.swglFillFramebuffer(w, C.c_float(d)),
                         lambda: np.ctypeslib.as_array(gpu_api.swglGetDepthPtr(), shape=(H, W)).copy(),
                         perturb=perturb, devices=devices, ours=True, options={"mip_lod": 1} if lod else None)
    finally:
        gpu_api.swglSetOption(b"mip_lod", 0)
        for name, values in [p for p in PERTURB if len(p) == 2]:       # back to the defaults for whoever comes next
            gpu_api.swglSetOption(name.encode(), {"host_mirror": 1, "fuse_clear": 1, "lean_prims": 1, "setup_big": 1, "jit": 1,
                                                  "overflow_pool": 1, "bin_cap": 256, "count_fragments": 1}.get(name, 0))
    err = gpu_api.swglGetLastError().decode()
    if err:
        return f"{what}: {err}"
    rf, rd = run_ops(reference.api, ops, lambda w, d: reference.lib.swglref_fill(w, C.c_float(d)),
                     lambda: np.ctypeslib.as_array(reference.lib.swglref_depth_ptr(), shape=(H, W)).copy())
    for i, (a, b) in enumerate(zip(gf, rf)):
        if not np.array_equal(a, b):
            return f"{what}: frame {i} of {len(gf)} differs in {int((a != b).sum())} pixels; ops {[o[0] for o in ops]}"
    cmp = O.compare(gf[-1], gd, rf[-1], rd)
    if cmp["color_mismatch"] or cmp["depth_mismatch"] or cmp["coverage_mismatch"]:
        return f"{what}: {cmp}"
    return ""


def compare_seed_in_ranks(gpu_api, reference, seed, ranks, band_rows=1, perturb=False):
    """The sequence once per sort-first rank (swglSetStripe: the rank rasterises and clears the bands it owns only), the
    ranks' bands stitched together like the peer / shared-mirror targets assemble them, against the reference's frames.
    Viewports stay inside the framebuffer rows (ranks of separate processes refuse folded draws)."""
    ops = make_ops(seed, inside=True)
    W, H = ops[0][3], ops[0][4]
    owner = ((np.arange(H) >> 5) // band_rows) % ranks
    frames, depth = None, None
    try:
        for r in range(ranks):
            gf, gd = run_ops(gpu_api, ops, lambda w, d: gpu_api.swglFillFramebuffer(w, C.c_float(d)),
                             lambda: np.ctypeslib.as_array(gpu_api.swglGetDepthPtr(), shape=(H, W)).copy(),
                             perturb=5000 + seed if perturb else None, ours=True, stripe=(r, ranks, band_rows))
            err = gpu_api.swglGetLastError().decode()
            if err:
                return f"seed {seed} rank {r}: {err}"
            if frames is None:
                frames, depth = [f.copy() for f in gf], gd.copy()
            for f, g in zip(frames, gf):
                f[owner == r] = g[owner == r]
            depth[owner == r] = gd[owner == r]
    finally:
        gpu_api.swglSetStripe(0, 1, 1)
        for name, values in [p for p in PERTURB if len(p) == 2]:
            gpu_api.swglSetOption(name.encode(), {"host_mirror": 1, "fuse_clear": 1, "lean_prims": 1, "setup_big": 1, "jit": 1,
                                                  "overflow_pool": 1, "bin_cap": 256, "count_fragments": 1}.get(name, 0))
    rf, rd = run_ops(reference.api, ops, lambda w, d: reference.lib.swglref_fill(w, C.c_float(d)),
                     lambda: np.ctypeslib.as_array(reference.lib.swglref_depth_ptr(), shape=(H, W)).copy())
    for i, (a, b) in enumerate(zip(frames, rf)):
        if not np.array_equal(a, b):
            return f"seed {seed}: stitched frame {i} of {len(frames)} differs in {int((a != b).sum())} pixels; ops {[o[0] for o in ops]}"
    cmp = O.compare(frames[-1], depth, rf[-1], rd)
    if cmp["color_mismatch"] or cmp["depth_mismatch"] or cmp["coverage_mismatch"]:
        return f"seed {seed}: {cmp}"
    return ""


@pytest.mark.parametrize("seed,ranks,band_rows", [(30, 2, 1), (31, 3, 1), (32, 2, 2), (33, 4, 1)])
def test_random_call_sequence_in_sort_first_ranks(gpu_api, reference, seed, ranks, band_rows):
    assert compare_seed_in_ranks(gpu_api, reference, seed, ranks, band_rows, perturb=seed % 2 == 1) == ""


@pytest.mark.parametrize("seed", range(12))
def test_random_call_sequence_matches_compiled_reference(gpu_api, reference, seed):
    assert compare_seed(gpu_api, reference, seed) == ""


@pytest.mark.parametrize("seed", range(12, 24))
def test_random_call_sequence_with_library_options_changing_underneath(gpu_api, reference, seed):
    """The same, with tuning options, waits, statistics and the other read-back calls slipped in between the
    reference's calls on the library's side: none of them may change a bit of any frame."""
    assert compare_seed(gpu_api, reference, seed, perturb=True) == ""


@pytest.mark.parametrize("seed,devices", [(24, 2), (25, 2), (26, 3), (27, 4), (239, 2), (635, 3)])
def test_random_call_sequence_on_a_device_group(gpu_api, reference, seed, devices, monkeypatch):
    """The same through swglSetDeviceCount (members wrap around the visible devices when there are fewer); viewports that
    leave the framebuffer rows are folded on the leader with every member's bands gathered there (group_draw_folded)."""
    monkeypatch.setenv("SWGL_B200_GROUP_EMULATE", "1")
    assert compare_seed(gpu_api, reference, seed, perturb=seed % 2 == 0, devices=devices) == ""


@pytest.fixture(scope="module")
def reference_lod():
    try:
        return O.Reference(defined_rsqrt=True)
    except Exception as e:                                  # pragma: no cover
        pytest.skip(f"oracle/_ref/libswgl_ref_lod.so not available: {e}")


@pytest.mark.parametrize("seed", range(1000, 1008))
def test_random_call_sequence_with_mip_levels(gpu_api, reference_lod, seed):
    """Sequences that build mip chains, with the library's "mip_lod" on, against the reference compiled with the
    defined rsqrt (the level of detail is undefined behaviour in the plain build)."""
    assert compare_seed(gpu_api, reference_lod, seed, lod=True) == ""
