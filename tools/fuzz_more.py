"""Many more seeds of tests/test_fuzz_gpu.py's scene generator against the C restatement, through the variants of
the warp rasteriser and of the set-up stage that the fixed test suite pins less densely:

    python tools/fuzz_more.py <first> <last>

round 2: tile heights 8 / 4 / 2, the thread-per-triangle and warp-per-triangle set-up (setup_big 0 / 1), the overflow
pool with a tiny list capacity (bin_cap 8: most lists overflow into the pool or force K to grow), record-less
primitives on / off, the software-pipelined set-up kernel, the CTA cross-check kernel; and for every fourth seed a
viewport pushed beyond the framebuffer rows (folded draw, compared with the compiled reference when it is there)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import swgl_b200
from oracle import pyoracle as O
from test_fuzz_gpu import _random_scene
from util import gpu_render

api = swgl_b200.load()
rest = O.Restatement()
try:
    ref = O.Reference()
except Exception:
    ref = None
VARIANTS = [{"raster_path": 0}, {"raster_path": 0, "lean_prims": 0}, {"raster_path": 2},
            {"raster_path": 3, "tile_rows": 8, "setup_big": 0}, {"raster_path": 3, "tile_rows": 4, "setup_big": 1},
            {"raster_path": 3, "tile_rows": 2}, {"raster_path": 3, "bin_cap": 8}, {"raster_path": 3, "bin_cap": 8, "overflow_pool": 0},
            {"raster_path": 3, "setup_pipelined": 1, "setup_big": 0}]
bad = 0
n = 0
folded = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    sc, rng = _random_scene(seed)
    fill = (int(rng.integers(0, 1 << 32)), float(rng.choice([0.0, 0.5, -1.0])))
    clear = bool(rng.random() < 0.7)
    rc, rd, rstats = rest.render(sc, clear=clear, fill=fill)
    for opts in VARIANTS:
        col, dep, stats, err = gpu_render(api, sc, indexed=sc.indices is not None, clear=clear, fill=fill, options=opts)
        cmp = O.compare(col, dep, rc, rd)
        ok = not err and cmp["coverage_mismatch"] == 0 and cmp["depth_mismatch"] == 0 and cmp["color_mismatch"] == 0 \
            and stats["tested"] == rstats["tested"] and stats["shaded"] == rstats["shaded"]
        n += 1
        if not ok:
            bad += 1
            print("MISMATCH", seed, sc.name, opts, err, cmp, stats["tested"], rstats["tested"])
    if ref is not None and seed % 4 == 0:
        # folded viewport: against the compiled reference (and the restatement's fragment counts)
        vp = sc.viewport or (0, 0, sc.width, sc.height)
        sc.viewport = (vp[0], vp[1] - int(rng.integers(1, 40)), vp[2], vp[3] + int(rng.integers(0, 60)))
        fc, fd = ref.render(sc, clear=clear, fill=fill)
        _, _, fstats = rest.render(sc, clear=clear, fill=fill)          # the restatement folds too; it also counts fragments
        col, dep, stats, err = gpu_render(api, sc, indexed=sc.indices is not None, clear=clear, fill=fill)
        cmp = O.compare(col, dep, fc, fd)
        if (stats["tested"], stats["shaded"]) != (fstats["tested"], fstats["shaded"]):
            cmp["coverage_mismatch"] += 1
            print("COUNTS folded", seed, stats["tested"], fstats["tested"], stats["shaded"], fstats["shaded"])
        n += 1
        outside = sc.viewport[1] < 0 or sc.viewport[1] + sc.viewport[3] > sc.height
        folded += int(api.swglGetOption(b"draws_folded") >= 1)
        if err or cmp["coverage_mismatch"] or cmp["depth_mismatch"] or cmp["color_mismatch"] or (outside and api.swglGetOption(b"draws_folded") < 1):
            bad += 1
            print("MISMATCH folded", seed, sc.name, sc.viewport, err, cmp)
print("seeds", sys.argv[1], sys.argv[2], "renders", n, "of which folded draws", folded, "mismatches", bad)
