"""GPU parity: the CUDA draw path (through the C ABI) against the CPU oracle, bit for bit.

Bar (BASELINE.json north_star): coverage and depth bit-exact, colour within 1/255 per channel
(asserted exact here, mismatches would be counted and reported by ``compare``).
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import gl as G, scenes as S

from util import assert_bit_exact, gpu_render

pytestmark = pytest.mark.gpu


def _small_scenes():
    out = [
        S.single_triangle(),
        S.random_triangles(),                                    # BASELINE config 1
        S.random_triangles(textured=True),
        S.random_triangles(2000, near_cross=True, alpha=None, centre_range=1.3, seed=7),
        S.grid_mesh(32, 320, 200),
        S.grid_mesh(32, 320, 200, textured=True),
        S.grid_mesh(40, 333, 211, alpha=0.5, use_matrix=True),   # width not a multiple of 4
    ]
    for vp, seed in [((37, 11, 301, 257), 99), ((10, 20, 600, 430), 5), ((5, 7, 600, 400), 6)]:
        sc = S.random_triangles(1500, near_cross=True, alpha=None, centre_range=1.3, seed=seed)
        sc.viewport = vp
        sc.name += f"_vp{vp[0]}"
        out.append(sc)
    return out


# 3: the warp rasteriser with the tile height it picks itself (2 rows for framebuffers this small); 38 / 34: pinned
# to 8 and 4 rows (the heights of large framebuffers and of sort-first ranks)
PATHS = {1: "pixel_owner", 2: "fragment_parallel", 3: "warp_tile", 38: "warp_tile_8_rows", 34: "warp_tile_4_rows"}


def _opts(path):
    return {"raster_path": 3, "tile_rows": path % 10} if path > 3 else {"raster_path": path}


@pytest.mark.parametrize("path", sorted(PATHS), ids=lambda p: PATHS[p])
@pytest.mark.parametrize("scene", _small_scenes(), ids=lambda s: s.name)
def test_matches_reference_and_restatement(gpu_api, restatement, reference, scene, path):
    col, dep, stats, err = gpu_render(gpu_api, scene, indexed=scene.indices is not None,
                                      options=_opts(path))
    assert err == "", err
    rc, rd, rstats = restatement.render(scene)
    assert_bit_exact(O.compare(col, dep, rc, rd), "vs restatement")
    fc, fd = reference.render(scene)
    assert_bit_exact(O.compare(col, dep, fc, fd), "vs compiled reference")
    assert stats["tested"] == rstats["tested"]
    assert stats["shaded"] == rstats["shaded"]
    assert stats["prims_out"] <= rstats["prims_out"]  # primitives with no rows are dropped early


def test_draw_arrays_equals_draw_elements(gpu_api):
    scene = S.grid_mesh(48, 400, 300)
    a = gpu_render(gpu_api, scene, indexed=True)
    b = gpu_render(gpu_api, scene, indexed=False)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_unfused_clear_equals_fused(gpu_api, restatement):
    scene = S.random_triangles(300, 320, 240, seed=3)
    scene.viewport = (16, 8, 280, 200)
    a = gpu_render(gpu_api, scene, fill=(0x11223344, 0.5), options={"fuse_clear": 1})
    b = gpu_render(gpu_api, scene, fill=(0x11223344, 0.5), options={"fuse_clear": 0})
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    rc, rd, _ = restatement.render(scene, fill=(0x11223344, 0.5))
    assert_bit_exact(O.compare(a[0], a[1], rc, rd))


def test_no_clear_blends_over_existing_frame(gpu_api, restatement):
    scene = S.random_triangles(200, 320, 240, seed=11, alpha=0.5)
    col, dep, _, err = gpu_render(gpu_api, scene, clear=False, fill=(0x80402010, 0.0))
    assert err == ""
    rc, rd, _ = restatement.render(scene, clear=False, fill=(0x80402010, 0.0))
    assert_bit_exact(O.compare(col, dep, rc, rd))


def test_two_draws_keep_order(gpu_api, restatement):
    scene = S.random_triangles(400, 320, 240, seed=21, alpha=0.5)
    col, dep, _, err = gpu_render(gpu_api, scene, draws=[(0, 600), (600, 600)])
    assert err == ""
    rc, rd, _ = restatement.render(scene)
    assert_bit_exact(O.compare(col, dep, rc, rd))


def test_long_tile_lists_and_span_pool_cuts(gpu_api, restatement):
    """Many large overlapping triangles: tile lists longer than one batch (256), more spans than
    the per-batch pool, and lists beyond the shared-memory sort capacity (2048)."""
    scene = S.random_triangles(6000, 256, 192, seed=77, extent=0.9, alpha=0.5)
    rc, rd, rstats = restatement.render(scene)
    for path in (1, 2, 3, 38, 34):
        col, dep, stats, err = gpu_render(gpu_api, scene, options=_opts(path))
        assert err == ""
        assert_bit_exact(O.compare(col, dep, rc, rd), PATHS[path])
        assert stats["tested"] == rstats["tested"] and stats["shaded"] == rstats["shaded"]


def test_depth_zero_reopens_pixels(gpu_api, restatement):
    """z == 0.0 stored as depth means "empty" again (swgl.c:3387): later fragments that would
    fail against the previous depth pass.  Exercises the late-shade path of the commit loop."""
    scene = S.random_triangles(1500, 320, 240, seed=31, extent=0.5, alpha=0.5)
    v = scene.vertices.reshape(-1, 3, 8)
    v[::3, :, 2] = 0.0          # every third triangle lies exactly on z = 0
    v[1::7, :, 2] = -0.25       # some negative depths too
    rc, rd, rstats = restatement.render(scene)
    for path in (1, 2, 3, 38, 34):
        col, dep, stats, err = gpu_render(gpu_api, scene, options=_opts(path))
        assert err == ""
        assert_bit_exact(O.compare(col, dep, rc, rd), PATHS[path])
        assert stats["shaded"] == rstats["shaded"]


def test_bin_overflow_grows_and_splits(gpu_api, restatement):
    """Per-tile lists have a fixed capacity K.  A draw that overflows it is dropped by the device,
    then re-issued with a larger K -- or, when the lists may not grow any further, split into
    consecutive sub-draws (which keep the submission order).  Result must not change."""
    scene = S.random_triangles(1200, 256, 192, seed=5, extent=0.6, alpha=0.5)
    rc, rd, rstats = restatement.render(scene)
    for path in (1, 2, 3, 38, 34):
        # K = 4: grows until the longest list fits
        col, dep, stats, err = gpu_render(gpu_api, scene, options={**_opts(path), "bin_cap": 4})
        assert err == ""
        assert_bit_exact(O.compare(col, dep, rc, rd), "grow " + PATHS[path])
        assert gpu_api.swglGetOption(b"bin_cap") > 4
        # K = 8 and lists limited to 16 entries per tile: the draw has to be split recursively
        ntiles = ((scene.width + 31) // 32) * ((scene.height + 31) // 32)
        col, dep, stats, err = gpu_render(gpu_api, scene, options={**_opts(path), "bin_cap": 8,
                                                                   "bin_limit_bytes": ntiles * 16 * 4})
        assert err == ""
        assert_bit_exact(O.compare(col, dep, rc, rd), "split " + PATHS[path])


def test_tall_triangles_use_band_entries(gpu_api, restatement):
    """Triangles taller than SWGL_SHORT_ROWS keep per-band walk states; shorter ones are re-walked."""
    scene = S.random_triangles(300, 512, 768, seed=8, extent=0.95, alpha=0.7, centre_range=0.8)
    rc, rd, rstats = restatement.render(scene)
    for path in (1, 2, 3, 38, 34):
        col, dep, stats, err = gpu_render(gpu_api, scene, options=_opts(path))
        assert err == "" and stats["bands"] > 0
        assert_bit_exact(O.compare(col, dep, rc, rd), PATHS[path])


def test_wide_and_tall_primitives_inside_a_fine_mesh(gpu_api, restatement):
    """A draw of small triangles binds its few tall (> 2 tile heights) or wide (> 64 columns) primitives
    inline in the set-up kernel (band entries, no k_bin_tall launch); all others have no primitive
    record at all.  Mix both kinds into one indexed draw, in the middle of the submission order."""
    mesh = S.grid_mesh(70, 512, 384, alpha=0.6)
    rng = np.random.default_rng(7)
    extra = []
    for k in range(12):
        cx, cy = rng.uniform(-0.7, 0.7, 2)
        if k % 3 == 0:      # wide and flat
            tri = [(cx - 0.9, cy, 0.5), (cx + 0.9, cy + 0.01, 0.4), (cx, cy + 0.03, 0.6)]
        elif k % 3 == 1:    # tall and thin
            tri = [(cx, cy - 0.9, 0.3), (cx + 0.02, cy + 0.9, 0.5), (cx - 0.02, cy + 0.5, 0.7)]
        else:               # big
            tri = [(cx - 0.5, cy - 0.4, 0.2), (cx + 0.6, cy - 0.3, 0.8), (cx, cy + 0.7, 0.5)]
        for (x, y, z) in tri:
            w = rng.uniform(1.0, 1.5)
            extra.append([x * w, y * w, z, w, *rng.random(3), 0.6])
    extra = np.asarray(extra, dtype=np.float32)
    base = len(mesh.vertices)
    mesh.vertices = np.ascontiguousarray(np.concatenate([mesh.vertices, extra]))
    idx = mesh.indices
    half = (len(idx) // 6) * 3
    mesh.indices = np.ascontiguousarray(np.concatenate(
        [idx[:half], np.arange(base, base + len(extra), dtype=np.uint32), idx[half:]]))
    mesh.name += "_mixed"
    rc, rd, rstats = restatement.render(mesh)
    for path in (3, 38, 34, 2):
        col, dep, stats, err = gpu_render(gpu_api, mesh, options=_opts(path))
        assert err == ""
        assert_bit_exact(O.compare(col, dep, rc, rd), PATHS[path])
        assert stats["tested"] == rstats["tested"] and stats["shaded"] == rstats["shaded"]
        assert stats["bands"] > 0


@pytest.mark.parametrize("width", [9000, 16000, 40000])
def test_wide_framebuffers_around_the_fast_division_domain(gpu_api, restatement, width):
    """The shared-reciprocal division is used for primitives whose snapped coordinates stay within
    2^14 (swgl_dev_math.cuh); wider framebuffers put some or all primitives outside and on the plain
    `/` path.  Both must give the reference's bits."""
    scene = S.random_triangles(400, width, 96, seed=width, extent=0.05, alpha=0.5, near_cross=True, centre_range=1.1)
    rc, rd, rstats = restatement.render(scene)
    col, dep, stats, err = gpu_render(gpu_api, scene, indexed=False)
    assert err == ""
    assert_bit_exact(O.compare(col, dep, rc, rd), scene.name)
    assert stats["tested"] == rstats["tested"] and stats["shaded"] == rstats["shaded"]


@pytest.mark.parametrize("deep", [3000, 100000])
def test_one_deep_tile_goes_through_the_overflow_pool(gpu_api, restatement, deep):
    """Localised overdraw: `deep` small blended triangles on one spot of an otherwise ordinary mesh.  The
    lists of the few tiles under the spot continue in the overflow pool; K and the per-tile lists stay as
    they are and the draw is not issued a second time (3 kernels), unless the pool itself has to grow once."""
    import copy
    mesh = S.grid_mesh(24, 320, 200, alpha=0.5)
    rng = np.random.default_rng(3)
    spot = np.empty((deep * 3, 8), np.float32)
    c = np.array([0.31, -0.17], np.float32)
    spot[:, 0:2] = c + rng.uniform(-0.02, 0.02, (deep * 3, 2)).astype(np.float32)
    spot[:, 2] = rng.uniform(0.1, 0.9, deep * 3)
    spot[:, 3] = 1.0
    spot[:, 4:7] = rng.uniform(0, 1, (deep * 3, 3))
    spot[:, 7] = 0.5
    sc = copy.copy(mesh)
    sc.vertices = np.concatenate([mesh.deindexed(), spot]).astype(np.float32)
    sc.indices = None
    sc.name = f"deep_tile_{deep}"
    rc, rd, rstats = restatement.render(sc)
    gpu_render(gpu_api, sc, indexed=False)                         # first frame: scratch reaches its steady size
    api = gpu_api
    pairs0, k0 = api.swglGetOption(b"pairs_bytes"), api.swglGetOption(b"bin_cap")
    n0 = api.swglGetOption(b"kernel_launches")
    api.glClear(3)
    api.glDrawArrays(G.GL_TRIANGLES, 0, len(sc.vertices))
    api.swglFinish()
    assert api.swglGetOption(b"kernel_launches") - n0 == 3         # vertex, set-up, raster: no second issue
    assert api.swglGetOption(b"bin_cap") == k0 == 256 and api.swglGetOption(b"pairs_bytes") == pairs0
    col, dep, stats, err = gpu_render(gpu_api, sc, indexed=False)
    assert err == "", err
    assert_bit_exact(O.compare(col, dep, rc, rd), sc.name)
    assert stats["tested"] == rstats["tested"] and stats["shaded"] == rstats["shaded"]
    # the same scene with the pool switched off: K grows for every tile instead, same bits
    col2, dep2, _, err2 = gpu_render(gpu_api, sc, indexed=False, options={"overflow_pool": 0})
    assert err2 == "" and np.array_equal(col, col2) and np.array_equal(dep.view(np.uint32), dep2.view(np.uint32))
    if deep >= 100000:
        assert gpu_api.swglGetOption(b"bin_cap") > 256
