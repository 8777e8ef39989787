"""Test infrastructure (tests/test_host_sanitizers.py): the call sequences of tests/test_api_fuzz_gpu.py -- plain, with
the library-only calls slipped in, with mip levels -- through the host layer built with AddressSanitizer / UBSan over
the stub device (SWGL_B200_LIB), which reads every byte a draw hands it.

    python host_sequences.py <first seed> <last seed>"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import swgl_b200
from swgl_b200 import gl as G
api = swgl_b200.load()        # SWGL_B200_LIB points at the sanitizer build of the host layer over the stub device
import test_api_fuzz_gpu as F
lo, hi = int(sys.argv[1]), int(sys.argv[2])
n = 0
for seed in range(lo, hi):
    for lod in (False, True):
        ops = F.make_ops(seed, lod=lod)
        W, H = ops[0][3], ops[0][4]
        F.run_ops(api, ops, lambda w, d: api.swglFillFramebuffer(w, C.c_float(d)),
                  lambda: np.ctypeslib.as_array(api.swglGetDepthPtr(), shape=(H, W)).copy(),
                  perturb=5000 + seed if seed % 2 else None, ours=True, options={"mip_lod": 1} if lod else None)
        n += 1
print("sequences through the host layer:", n, flush=True)
os._exit(0)
