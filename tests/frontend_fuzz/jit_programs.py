"""Test infrastructure (tests/test_host_sanitizers.py): every shader case and a range of random programs through
swglPrecompileProgram of the sanitizer build -- the IR -> CUDA C++ generator of swgl_jit.cpp and an NVRTC compilation for
sm_100a, no device.

    python jit_programs.py <first random seed> <last random seed>"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import swgl_b200
from swgl_b200 import gl as G
from shader_cases import CASES
from shader_fuzz_gen import make
from test_jit import _compile
api = swgl_b200.load()
api.glInit(64, 48)
progs = [(vs, fs) for vs, fs, _ in CASES.values()] + [make(k)[:2] for k in range(int(sys.argv[1]), int(sys.argv[2]))]
ok = 0
for vs, fs in progs:
    api.swglGetLastError()
    rc = _compile(api, vs, fs)
    err = api.swglGetLastError().decode()
    if rc != 0:
        print("precompile failed:", err[:300]); os._exit(3)
    ok += 1
print("programs generated and compiled:", ok, flush=True)
os._exit(0)
