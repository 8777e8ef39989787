"""CPU: the C host layer (swgl_host.c + swgl_glsl.c) under AddressSanitizer + UndefinedBehaviorSanitizer, over a stand-in for
the CUDA device layer that reads every byte a draw hands it (tests/frontend_fuzz/stub_dev.c): a pointer kept past a free, a
size that does not match its allocation or an index outside an object table shows here, on a box without a GPU.

Two drivers: the call sequences of the GPU fuzz (buffers re-specified in place, texture objects re-bound and re-specified,
mip chains appended, options changing under way), and tens of thousands of calls whose arguments make no sense."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "swgl_b200", "csrc")
FUZZ = os.path.join(ROOT, "tests", "frontend_fuzz")


@pytest.fixture(scope="module")
def sanitized_host(tmp_path_factory):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    so = str(tmp_path_factory.mktemp("asan") / "libswgl_host_asan.so")
    build = subprocess.run([gcc, "-std=gnu11", "-O1", "-g", "-fPIC", "-shared", "-w", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", CSRC, os.path.join(CSRC, "swgl_host.c"),
                            os.path.join(CSRC, "swgl_glsl.c"), os.path.join(FUZZ, "stub_dev.c"), "-o", so, "-lm"], capture_output=True, text=True)
    runtime = subprocess.run([gcc, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if build.returncode != 0 or not os.path.isabs(runtime) or not os.path.exists(runtime):
        pytest.skip("sanitizer runtime not available: " + build.stderr[-300:])
    return dict(os.environ, LD_PRELOAD=runtime, SWGL_B200_LIB=so, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1",
                UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")


def test_call_sequences_through_the_sanitized_host_layer(sanitized_host):
    run = subprocess.run([sys.executable, os.path.join(FUZZ, "host_sequences.py"), "0", "60"], env=sanitized_host, capture_output=True, timeout=900)
    assert run.returncode == 0 and b"sequences through the host layer: 120" in run.stdout, (run.stdout[-500:], run.stderr[-3000:])


@pytest.mark.parametrize("seed", [1, 2])
def test_nonsense_arguments_through_the_sanitized_host_layer(sanitized_host, seed):
    run = subprocess.run([sys.executable, os.path.join(FUZZ, "hostile_calls.py"), str(seed), "8000"], env=sanitized_host, capture_output=True, timeout=900)
    assert run.returncode == 0 and b"hostile calls: 8000" in run.stdout, (run.stdout[-500:], run.stderr[-3000:])
