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


def test_code_generator_of_the_run_time_compiler_under_the_sanitizers(sanitized_host, tmp_path):
    """swgl_jit.cpp (IR -> CUDA C++, NVRTC for sm_100a, no device) built with the sanitizers as well: the twelve shader
    cases and ten random programs are generated and compiled."""
    from swgl_b200 import build as B
    gxx, gcc = shutil.which("g++"), shutil.which("gcc")
    cuda = os.path.dirname(os.path.dirname(B.NVCC))
    if not gxx or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("no g++ / CUDA headers")
    B._write_jit_sources(str(tmp_path / "swgl_jit_sources.inc"))
    san = ["-O1", "-g", "-fPIC", "-w", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", str(tmp_path), "-I", os.path.join(cuda, "include")]
    objs = []
    for cc, std, src, extra in ((gxx, "-std=c++17", os.path.join(CSRC, "swgl_jit.cpp"), []), (gcc, "-std=gnu11", os.path.join(CSRC, "swgl_host.c"), ["-ffp-contract=off"]),
                                (gcc, "-std=gnu11", os.path.join(CSRC, "swgl_glsl.c"), ["-ffp-contract=off"]), (gcc, "-std=gnu11", os.path.join(FUZZ, "stub_dev.c"), ["-DSTUB_WITH_JIT"])):
        obj = str(tmp_path / (os.path.basename(src) + ".o"))
        r = subprocess.run([cc, std] + san + extra + inc + ["-c", src, "-o", obj], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        objs.append(obj)
    so = str(tmp_path / "libswgl_host_jit_asan.so")
    r = subprocess.run([gxx, "-shared", "-fsanitize=address,undefined", "-o", so] + objs + ["-L" + os.path.join(cuda, "lib64"), "-Wl,-rpath," + os.path.join(cuda, "lib64"),
                                                                                           "-lcudart", "-ldl", "-lpthread", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(sanitized_host, SWGL_B200_LIB=so, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:protect_shadow_gap=0")
    run = subprocess.run([sys.executable, os.path.join(FUZZ, "jit_programs.py"), "0", "10"], env=env, capture_output=True, timeout=1200)
    if b"libnvrtc" in run.stdout + run.stderr and run.returncode == 3:
        pytest.skip("NVRTC not available")
    assert run.returncode == 0 and b"programs generated and compiled: 22" in run.stdout, (run.stdout[-800:], run.stderr[-3000:])
