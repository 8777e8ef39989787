"""One-off: many more seeds of tests/test_fuzz_gpu.py's scene generator through the default raster
path (and lean_prims on/off), against the C restatement.  python tools/fuzz_more.py <first> <last>"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import swgl_b200
from oracle import pyoracle as O
from test_fuzz_gpu import _random_scene
from util import gpu_render

api = swgl_b200.load()
rest = O.Restatement()
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    sc, rng = _random_scene(seed)
    fill = (int(rng.integers(0, 1 << 32)), float(rng.choice([0.0, 0.5, -1.0])))
    clear = bool(rng.random() < 0.7)
    rc, rd, rstats = rest.render(sc, clear=clear, fill=fill)
    for opts in ({"raster_path": 0}, {"raster_path": 0, "lean_prims": 0}, {"raster_path": 2}):
        col, dep, stats, err = gpu_render(api, sc, indexed=sc.indices is not None, clear=clear, fill=fill, options=opts)
        cmp = O.compare(col, dep, rc, rd)
        ok = not err and cmp["coverage_mismatch"] == 0 and cmp["depth_mismatch"] == 0 and cmp["color_mismatch"] == 0 \
            and stats["tested"] == rstats["tested"] and stats["shaded"] == rstats["shaded"]
        if not ok:
            bad += 1
            print("MISMATCH", seed, sc.name, opts, err, cmp, stats["tested"], rstats["tested"])
print("seeds", sys.argv[1], sys.argv[2], "mismatches", bad)
