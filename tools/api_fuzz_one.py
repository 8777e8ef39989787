"""One seed of the API-sequence fuzz on the GPU library only (for compute-sanitizer): python tools/api_fuzz_one.py <seed>"""
import ctypes as C
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import swgl_b200
import test_api_fuzz_gpu as F

api = swgl_b200.load()
ops = F.make_ops(int(sys.argv[1]))
F.run_ops(api, ops, lambda w, d: api.swglFillFramebuffer(w, C.c_float(d)),
          lambda: np.ctypeslib.as_array(api.swglGetDepthPtr(), shape=(ops[0][4], ops[0][3])).copy())
print("error:", api.swglGetLastError().decode())
