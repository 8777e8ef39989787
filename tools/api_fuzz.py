"""Many more seeds of tests/test_api_fuzz_gpu.py: random sequences of the reference's API calls issued to
libswgl_b200.so and to the compiled reference, every frame read on the way and the final colour / depth compared.

    python tools/api_fuzz.py <first> <last> [perturb] [devices N] [lod] [ranks N]

perturb: library-only calls (tuning options, waits, statistics, the other read-back calls) slipped in between;
lod: mip_lod = 1 against the reference built with the defined rsqrt;
ranks N: the sequence once per sort-first rank (swglSetStripe), bands stitched, viewports inside the framebuffer;
devices N: through swglSetDeviceCount(N) (SWGL_B200_GROUP_EMULATE=1 wraps members around the visible devices)."""
import os
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import swgl_b200
from oracle import pyoracle as O
import test_api_fuzz_gpu as F

api = swgl_b200.load()
lod = "lod" in sys.argv
ref = O.Reference(defined_rsqrt=lod)
perturb = "perturb" in sys.argv
devices = int(sys.argv[sys.argv.index("devices") + 1]) if "devices" in sys.argv else 1
if devices > 1:
    os.environ.setdefault("SWGL_B200_GROUP_EMULATE", "1")
bad = n = draws = folded = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    if "ranks" in sys.argv:
        nr = int(sys.argv[sys.argv.index("ranks") + 1])
        msg = F.compare_seed_in_ranks(api, ref, seed, nr, band_rows=1 + seed % 2, perturb=perturb)
    else:
        msg = F.compare_seed(api, ref, seed, perturb=perturb, devices=devices, lod=lod)
    n += 1
    draws += sum(1 for o in F.make_ops(seed) if o[0] in ("draw", "points"))
    folded += int(api.swglGetOption(b"draws_folded"))
    if msg:
        bad += 1
        print("MISMATCH", msg)
print("api sequences", sys.argv[1], sys.argv[2], "perturb" if perturb else "", "lod" if lod else "", "devices", devices, "run", n, "draw calls", draws, "of which folded", folded, "mismatches", bad,
      "jit compiles", api.swglGetOption(b"jit_compiles"))
