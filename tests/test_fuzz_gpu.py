"""GPU: randomised scenes (sizes, viewports, alpha, near-plane crossings, textures, matrices, draw
splitting) through both raster kernels against the C restatement -- bit-exact, fragment counts too."""
import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import scenes as S

from util import assert_bit_exact, gpu_render

pytestmark = pytest.mark.gpu


def _random_scene(seed):
    rng = np.random.default_rng(1000 + seed)
    w = int(rng.choice([64, 97, 160, 256, 333, 512]))
    h = int(rng.choice([48, 75, 128, 200, 211, 384]))
    kind = rng.integers(0, 4)
    if kind == 0:
        sc = S.random_triangles(int(rng.integers(50, 1500)), w, h, seed=int(rng.integers(1, 1 << 30)),
                                extent=float(rng.choice([0.05, 0.2, 0.5, 1.2])),
                                alpha=None if rng.random() < 0.5 else float(rng.choice([1.0, 0.5, 0.1])),
                                near_cross=bool(rng.random() < 0.5), centre_range=float(rng.choice([0.7, 1.0, 1.4])),
                                textured=bool(rng.random() < 0.3))
    elif kind == 1:
        sc = S.grid_mesh(int(rng.integers(4, 90)), w, h, seed=int(rng.integers(1, 1 << 30)),
                         alpha=float(rng.choice([1.0, 0.5])), use_matrix=bool(rng.random() < 0.5))
    elif kind == 2:
        sc = S.grid_mesh(int(rng.integers(4, 60)), w, h, seed=int(rng.integers(1, 1 << 30)), textured=True)
        sc.tex_wrap = str(rng.choice(["repeat", "clamp"]))
    else:
        sc = S.grid_mesh(int(rng.integers(3, 30)), w, h, seed=int(rng.integers(1, 1 << 30)), layers=int(rng.integers(2, 6)),
                         alpha=float(rng.choice([1.0, 0.3])))
    if rng.random() < 0.5:
        vx, vy = int(rng.integers(0, w // 3)), int(rng.integers(0, h // 3))
        sc.viewport = (vx - (int(rng.integers(0, 40)) if rng.random() < 0.3 else 0), vy,
                       int(rng.integers(8, w - vx + (30 if rng.random() < 0.3 else 0) + 1)), int(rng.integers(8, h - vy + 1)))
    sc.clear_color = tuple(float(x) for x in rng.random(4))
    return sc, rng


@pytest.mark.parametrize("seed", range(24))
def test_random_scene_matches_restatement(gpu_api, restatement, seed):
    sc, rng = _random_scene(seed)
    fill = (int(rng.integers(0, 1 << 32)), float(rng.choice([0.0, 0.5, -1.0])))
    clear = bool(rng.random() < 0.7)
    rc, rd, rstats = restatement.render(sc, clear=clear, fill=fill)
    n = len(sc.indices) if sc.indices is not None else len(sc.vertices)
    cut = 3 * int(rng.integers(0, n // 3 + 1))
    for path, rows in ((1, 0), (2, 0), (3, 0), (3, 8), (3, 4)):       # warp rasteriser: its own choice of tile height (2 here), 8 and 4 rows
        draws = None if path == 1 else [(0, cut), (cut, n - cut)]       # also split the draw in two
        col, dep, stats, err = gpu_render(gpu_api, sc, indexed=sc.indices is not None, clear=clear, fill=fill,
                                          options={"raster_path": path, "tile_rows": rows, "fuse_clear": int(rng.integers(0, 2))}, draws=draws)
        assert err == "", err
        assert_bit_exact(O.compare(col, dep, rc, rd), f"{sc.name} path {path}")
