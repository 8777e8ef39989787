"""Random shader pairs (tests/shader_fuzz_gen.py) through the run-time compiled kernels, the on-device interpreter and
the CTA cross-check rasteriser against the COMPILED reference, which interprets the same source strings:

    python tools/shader_fuzz.py <first> <last>

Needs a GPU and oracle/_ref/libswgl_ref.so.  Prints one line per mismatch and a summary."""
import sys
import time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import swgl_b200
from oracle import pyoracle as O
from swgl_b200 import scenes as S
from shader_fuzz_gen import make
from util import gpu_render

api = swgl_b200.load()
ref = O.Reference()
VARIANTS = [{"raster_path": 3, "jit": 1}, {"raster_path": 3, "jit": 0}, {"raster_path": 2, "jit": 1}]
bad = n = 0
t0 = time.time()
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    vs, fs, uniforms = make(seed)
    rng = np.random.default_rng(seed)
    scene = S.random_triangles(int(rng.integers(100, 400)), 256, 192, seed=int(rng.integers(1, 1 << 30)),
                               alpha=None if rng.random() < 0.5 else 0.6, near_cross=bool(rng.random() < 0.3))
    scene.vs, scene.fs, scene.uniforms = vs, fs, uniforms
    fc, fd = ref.render(scene)
    for opts in VARIANTS:
        col, dep, stats, err = gpu_render(api, scene, options=opts)
        cmp = O.compare(col, dep, fc, fd)
        kinds = (api.swglGetOption(b"last_vs_kind"), api.swglGetOption(b"last_fs_kind"))
        n += 1
        wrong_kind = (opts["jit"] == 1 and opts["raster_path"] == 3 and 3 not in kinds) or (opts["jit"] == 0 and 3 in kinds)
        if err or wrong_kind or cmp["coverage_mismatch"] or cmp["depth_mismatch"] or cmp["color_mismatch"]:
            bad += 1
            print("MISMATCH", seed, opts, kinds, err, cmp)
            print(vs); print(fs); print(uniforms)
print("shader seeds", sys.argv[1], sys.argv[2], "renders", n, "mismatches", bad, "jit compiles", api.swglGetOption(b"jit_compiles"),
      "compile s", round(api.swglGetOption(b"jit_compile_us_total") * 1e-6, 1), "wall s", round(time.time() - t0, 1))
