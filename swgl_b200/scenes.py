"""Synthetic scene generators for the swgl draw-call path.

Every scene is a plain bundle of *bytes* (vertex array, optional u32 index array, optional
RGBA8 texture) plus the GLSL-subset shader pair, so exactly the same inputs can be fed to the
CUDA library, to the C restatement and to the compiled reference.  The PRNG is the 32-bit LCG
named in the survey: ``s = s*1664525 + 1013904223``, ``rnd() = (s >> 8) / 2**24``; it is
evaluated for the whole scene at once with wrapping uint32 cumulative products (jump-ahead),
so 8K scenes with millions of vertices are generated in well under a second.

All arithmetic that shapes the vertices is float32, left to right, so the arrays are
reproducible bit for bit on any host.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

LCG_A = np.uint32(1664525)
LCG_C = np.uint32(1013904223)

# Shader pairs accepted by the reference (layout declarations carry no
# `in` keyword; newlines/tabs are deleted by the front-end without inserting a space).
VS_PASSTHROUGH = (
    "layout (location = 0) vec4 aPos;\n"
    "layout (location = 1) vec4 aCol;\n"
    "out vec4 vCol;\n"
    "void main()\n{\n"
    "gl_Position = aPos;\n"
    "vCol = aCol;\n"
    "}\n"
)
VS_MATRIX = (
    "layout (location = 0) vec4 aPos;\n"
    "layout (location = 1) vec4 aCol;\n"
    "uniform mat4 uM;\n"
    "out vec4 vCol;\n"
    "void main()\n{\n"
    "gl_Position = uM * aPos;\n"
    "vCol = aCol;\n"
    "}\n"
)
FS_COLOR = (
    "in vec4 vCol;\n"
    "out vec4 FragColor;\n"
    "void main()\n{\n"
    "FragColor = vCol;\n"
    "}\n"
)
VS_TEX = (
    "layout (location = 0) vec4 aPos;\n"
    "layout (location = 1) vec2 aUV;\n"
    "uniform mat4 uM;\n"
    "out vec2 vUV;\n"
    "void main()\n{\n"
    "gl_Position = uM * aPos;\n"
    "vUV = aUV;\n"
    "}\n"
)
FS_TEX = (
    "in vec2 vUV;\n"
    "uniform sampler2D uTex;\n"
    "out vec4 FragColor;\n"
    "void main()\n{\n"
    "FragColor = texture(uTex, vUV);\n"
    "}\n"
)
# K2 of SURVEY.md appendix C: texture addressed with the xy of a vec4 varying.
FS_TEX_SWZ = (
    "in vec4 vCol;\n"
    "uniform sampler2D uTex;\n"
    "out vec4 FragColor;\n"
    "void main()\n{\n"
    "FragColor = texture(uTex, vCol.xy);\n"
    "}\n"
)


def lcg_stream(seed: int, n: int) -> np.ndarray:
    """Return ``rnd()`` for steps 1..n of the LCG as float32 in [0, 1)."""
    if n == 0:
        return np.zeros(0, np.float32)
    with np.errstate(over="ignore"):
        a_pow = np.cumprod(np.full(n, LCG_A, np.uint32), dtype=np.uint32)  # a^1 .. a^n
        geo = np.empty(n, np.uint32)  # 1 + a + ... + a^(k-1) for k = 1..n
        geo[0] = 1
        if n > 1:
            geo[1:] = a_pow[:-1]
        geo = np.cumsum(geo, dtype=np.uint32)
        s = a_pow * np.uint32(seed & 0xFFFFFFFF) + LCG_C * geo
    return (s >> np.uint32(8)).astype(np.float32) / np.float32(16777216.0)


@dataclass
class Scene:
    name: str
    width: int
    height: int
    vertices: np.ndarray  # float32 [V, floats_per_vertex], AoS
    attribs: list  # [(location, n_floats, byte_offset)]
    vs: str
    fs: str
    indices: Optional[np.ndarray] = None  # uint32 [I] or None (glDrawArrays)
    texture: Optional[np.ndarray] = None  # uint8 [H, W, 4] or None
    tex_wrap: str = "repeat"
    matrix: Optional[np.ndarray] = None  # float32 [16] handed to glUniformMatrix4fv(GL_FALSE)
    viewport: Optional[tuple] = None  # (x, y, w, h); default = full framebuffer
    clear_color: tuple = (0.0, 0.0, 0.0, 1.0)
    uniforms: dict = field(default_factory=dict)  # name -> ("1f"|"2f"|"3f"|"4f"|"1i"|"m2"|"m3"|"m4", values)
    meta: dict = field(default_factory=dict)

    @property
    def stride(self) -> int:
        return int(self.vertices.shape[1]) * 4

    @property
    def n_triangles(self) -> int:
        n = len(self.indices) if self.indices is not None else len(self.vertices)
        return n // 3

    def deindexed(self) -> np.ndarray:
        """Vertex stream as glDrawArrays would consume it (what the reference draws)."""
        if self.indices is None:
            return self.vertices
        return np.ascontiguousarray(self.vertices[self.indices])

    def algorithmic_bytes(self) -> int:
        """V*stride + I*4 + texture + W*H*8 (frame starts with a clear)."""
        b = self.vertices.nbytes + self.width * self.height * 8
        if self.indices is not None:
            b += self.indices.nbytes
        if self.texture is not None:
            b += self.texture.nbytes
        return int(b)


def _u(r: np.ndarray, lo: float, hi: float) -> np.ndarray:
    lo32, hi32 = np.float32(lo), np.float32(hi)
    return lo32 + (hi32 - lo32) * r


def single_triangle(width: int = 64, height: int = 48) -> Scene:
    """K0: one RGB triangle."""
    v = np.array(
        [
            [-0.8, -0.8, 0.5, 1.0, 1.0, 0.0, 0.0, 1.0],
            [0.8, -0.8, 0.5, 1.0, 0.0, 1.0, 0.0, 1.0],
            [0.0, 0.8, 0.5, 1.0, 0.0, 0.0, 1.0, 1.0],
        ],
        np.float32,
    )
    return Scene("k0_single", width, height, v, [(0, 4, 0), (1, 4, 16)], VS_PASSTHROUGH, FS_COLOR)


def random_triangles(
    n_tris: int = 1000,
    width: int = 640,
    height: int = 480,
    seed: int = 12345,
    extent: float = 0.3,
    alpha: Optional[float] = 1.0,
    near_cross: bool = False,
    centre_range: float = 1.0,
    textured: bool = False,
) -> Scene:
    """C1: random per-vertex-colour triangles (glDrawArrays).

    Draw order per triangle: cx, cy, z, then per vertex w, dx, dy, r, g, b (21 draws).
    ``near_cross`` spreads clip z over [-1.3w, 1.1w] so about a third of the triangles cross the
    near plane; ``alpha=None`` draws alpha (and lets colours leave [0,1]) from the stream too.
    """
    per_tri = 3 + 3 * (7 if alpha is None else 6)
    r = lcg_stream(seed, n_tris * per_tri).reshape(n_tris, per_tri)
    cx = _u(r[:, 0], -centre_range, centre_range)
    cy = _u(r[:, 1], -centre_range, centre_range)
    z = _u(r[:, 2], 0.01, 0.99)
    k = 7 if alpha is None else 6
    verts = np.empty((n_tris, 3, 8), np.float32)
    for j in range(3):
        b = 3 + j * k
        w = _u(r[:, b + 0], 1.0, 2.0) if not near_cross else _u(r[:, b + 0], 0.5, 2.0)
        dx = _u(r[:, b + 1], -extent, extent)
        dy = _u(r[:, b + 2], -extent, extent)
        verts[:, j, 0] = (cx + dx) * w
        verts[:, j, 1] = (cy + dy) * w
        if near_cross:
            verts[:, j, 2] = _u(r[:, b + 3], -1.3, 1.1) * w
        else:
            verts[:, j, 2] = z
        verts[:, j, 3] = w
        if alpha is None:
            verts[:, j, 4:7] = _u(r[:, b + 3 : b + 6], -0.25, 1.25)
            verts[:, j, 7] = _u(r[:, b + 6], -0.25, 1.25)
        else:
            verts[:, j, 4:7] = r[:, b + 3 : b + 6]
            verts[:, j, 7] = np.float32(alpha)
    sc = Scene(
        f"random_{n_tris}_{width}x{height}_s{seed}",
        width,
        height,
        verts.reshape(-1, 8),
        [(0, 4, 0), (1, 4, 16)],
        VS_PASSTHROUGH,
        FS_COLOR,
    )
    if textured:
        sc.fs = FS_TEX_SWZ
        sc.texture = checker_texture(256)
        sc.name += "_tex"
    return sc


def checker_texture(n: int = 256) -> np.ndarray:
    """K2 texture: (x, y, ((x/16 + y/16) & 1) * 255, 255)."""
    y, x = np.mgrid[0:n, 0:n]
    t = np.empty((n, n, 4), np.uint8)
    t[..., 0] = x & 255
    t[..., 1] = y & 255
    t[..., 2] = (((x // 16) + (y // 16)) & 1) * 255
    t[..., 3] = 255
    return t


def lcg_texture(n: int = 1024, seed: int = 777) -> np.ndarray:
    """C3 texture: RGBA8 noise from the LCG (top byte of each state)."""
    r = lcg_stream(seed, n * n * 4)
    return (r * np.float32(256.0)).astype(np.uint8).reshape(n, n, 4)


def grid_mesh(
    grid: int,
    width: int,
    height: int,
    seed: int = 12345,
    alpha: float = 1.0,
    textured: bool = False,
    use_matrix: bool = False,
    layers: int = 1,
) -> Scene:
    """C2..C5: indexed jittered grid of ``grid**2`` quads, two triangles per quad.

    Per-vertex draw order: w, jx, jy, z, r, g, b.  x = -0.98 + 1.96*i/G + (jx-0.5)*0.6/G, same
    for y; positions are pre-multiplied by w so the perspective divide returns them.
    """
    g1 = grid + 1
    nv = g1 * g1 * layers
    r = lcg_stream(seed, nv * 7).reshape(nv, 7)
    jj, ii = np.mgrid[0:g1, 0:g1]
    ii = np.tile(ii.reshape(-1), layers).astype(np.float32)
    jj = np.tile(jj.reshape(-1), layers).astype(np.float32)
    G = np.float32(grid)
    w = _u(r[:, 0], 1.0, 1.5)
    x = np.float32(-0.98) + np.float32(1.96) * ii / G + (r[:, 1] - np.float32(0.5)) * np.float32(0.6) / G
    y = np.float32(-0.98) + np.float32(1.96) * jj / G + (r[:, 2] - np.float32(0.5)) * np.float32(0.6) / G
    z = _u(r[:, 3], 0.1, 0.9)
    if textured:
        verts = np.empty((nv, 6), np.float32)
        verts[:, 4] = ii / G * np.float32(4.0)  # uv in [0,4]^2 exercises GL_REPEAT
        verts[:, 5] = jj / G * np.float32(4.0)
        attribs = [(0, 4, 0), (1, 2, 16)]
    else:
        verts = np.empty((nv, 8), np.float32)
        verts[:, 4:7] = r[:, 4:7]
        verts[:, 7] = np.float32(alpha)
        attribs = [(0, 4, 0), (1, 4, 16)]
    verts[:, 0] = x * w
    verts[:, 1] = y * w
    verts[:, 2] = z
    verts[:, 3] = w
    qj, qi = np.mgrid[0:grid, 0:grid]
    a = (qj * g1 + qi).reshape(-1).astype(np.uint32)
    b = a + 1
    c = a + g1
    d = c + 1
    one = np.stack([a, b, c, b, d, c], axis=1).reshape(-1)
    idx = np.concatenate([one + np.uint32(l * g1 * g1) for l in range(layers)]).astype(np.uint32)
    matrix = None
    vs, fs = VS_PASSTHROUGH, FS_COLOR
    if textured:
        vs, fs = VS_TEX, FS_TEX
        use_matrix = True
    if use_matrix:
        # value[] handed to glUniformMatrix4fv(loc, 1, GL_FALSE, value).  With the reference's
        # transposing store and its mat4 load quirk (swgl.c:2231-2235, 3911-3926) the w row
        # reads value[10], value[7], value[11], value[15]; this choice keeps w' == w.
        matrix = np.array(
            [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0.0, 0, 0, 0, 0, 1.0], np.float32
        )
        # z' = value[2]*x + value[6]*y + value[10]*z + value[14]*w : make it 0.25*w + small xy tilt
        matrix[14] = 0.25
        matrix[2] = 0.05
        matrix[6] = -0.03
        if not textured:
            vs = VS_MATRIX
    sc = Scene(
        f"grid_{grid}_{width}x{height}_s{seed}" + ("_tex" if textured else "") + (f"_a{alpha}" if alpha != 1.0 else ""),
        width,
        height,
        verts,
        attribs,
        vs,
        fs,
        indices=idx,
        matrix=matrix,
    )
    if textured:
        sc.texture = lcg_texture(1024)
    return sc


# ---- the five BASELINE.json configurations -------------------------------------------------

def config(n: int) -> Scene:
    if n == 1:
        return random_triangles(1000, 640, 480, seed=12345)
    if n == 2:
        return grid_mesh(224, 1920, 1080)
    if n == 3:
        return grid_mesh(224, 1920, 1080, textured=True)
    if n == 4:
        return grid_mesh(708, 3840, 2160)
    if n == 5:
        return grid_mesh(1416, 7680, 4320, alpha=0.5)
    raise ValueError(f"no BASELINE config {n}")


CONFIG_NAMES = {
    1: "C1: 1,000 random per-vertex-colour triangles, 640x480, glDrawArrays",
    2: "C2: 1080p 100,352-triangle indexed grid, Gouraud varyings, depth test, opaque",
    3: "C3: 1080p textured grid (1024^2 RGBA8, nearest, REPEAT), perspective-correct UV",
    4: "C4: 4K 1,002,528 small triangles, depth test, indexed draw",
    5: "C5: 8K 4,010,112 triangles, alpha 0.5 blending over depth, indexed draw",
}
