"""CPU: the GLSL-subset front end under AddressSanitizer + UndefinedBehaviorSanitizer on mutated shader sources.

The reference dereferences NULL on most malformed sources; the front end here is meant to report instead, so it is held to
"no memory error, no undefined behaviour, no leak" on arbitrary text: every source of tests/shader_cases.py and thirty
random programs of tests/shader_fuzz_gen.py, mutated (pieces deleted, moved, inserted, truncated, stray bytes).
A first run of this found a real defect: the scanners' "not found" value was INT_MAX and `cursor = position + 1`
overflowed for a declaration without its semicolon at the end of a source (swgl_glsl.c: BIG)."""
import os
import random
import shutil
import struct
import subprocess

import pytest

from shader_cases import CASES
from shader_fuzz_gen import make

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "swgl_b200", "csrc")

FRAGMENTS = ["vec4", "vec3", "vec2", "mat4", "mat3", "mat2", "(", ")", ";", ",", ".", "=", "*", "+", "-", "/", "{", "}", "uniform ", "in ", "out ",
             "layout (location = 0) ", "layout (location = 99999) ", "texture(", "sin(", "cos(", "tan(", "min(", "max(", " ", "\n", "\t", "1.0", "0",
             ".xyzw", ".stpq", "float", "int", "sampler2D", "main", "void ", "gl_Position", "<", ">", "==", "9999999999999999999999", "1e38", ".x.y",
             "((((", "))))", "vec4(1,2,3,4,5,6,7,8,9)", "vec4()", "float(", "int(", "-.", "..", "-----", "a" * 300, "(" * 200, "1" * 400, "vec4 " * 50]


def mutants(seed, n):
    rnd = random.Random(seed)
    pool = [src for vs, fs, _ in CASES.values() for src in (vs, fs)]
    for k in range(30):
        vs, fs, _ = make(k)
        pool += [vs, fs]
    out = bytearray()
    for _ in range(n):
        s = pool[rnd.randrange(len(pool))]
        for _ in range(rnd.randint(1, 6)):
            r, p = rnd.random(), rnd.randrange(len(s) + 1)
            if r < 0.3 and len(s) > 2:
                s = s[:p] + s[min(len(s), p + rnd.randint(1, 20)):]
            elif r < 0.7:
                s = s[:p] + FRAGMENTS[rnd.randrange(len(FRAGMENTS))] + s[p:]
            elif r < 0.85 and len(s) > 2:
                q = rnd.randrange(len(s))
                a, b = min(p, q), max(p, q)
                s = s[:a] + s[b:] + s[a:b]
            elif r < 0.95:
                s = s[:p] + chr(rnd.randint(1, 255)) + s[p:]
            else:
                s = s[:p]
        b = s.encode("latin-1", "replace").replace(b"\x00", b" ")
        out += struct.pack("<I", len(b)) + b
    return bytes(out)


def test_front_end_is_clean_under_asan_and_ubsan_on_mutated_sources(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = str(tmp_path / "harness")
    build = subprocess.run([gcc, "-std=gnu11", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-ffp-contract=off",
                            "-I", CSRC, os.path.join(ROOT, "tests", "frontend_fuzz", "harness.c"), os.path.join(CSRC, "swgl_glsl.c"),
                            "-o", exe, "-lm"], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + build.stderr[-300:])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    n = 8000
    run = subprocess.run([exe], input=mutants(2024, n), capture_output=True, env=env, timeout=600)
    assert run.returncode == 0, run.stderr.decode(errors="replace")[-2000:]
    assert f"compiled {n} records".encode() in run.stdout
