"""The drop-in claim, literally: a plain C application (two translation units that both include
swgl.h) compiled with gcc and linked against libswgl_b200.so instead of swgl.c.
CPU: it compiles and links.  GPU: it renders known-answer scene K0 with the reference's result."""
import os
import subprocess

import pytest

from swgl_b200._lib import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.dirname(LIB_PATH)


def _build(tmp_path, name="k0", sources=("k0_triangle.c", "k0_hash.c")):
    exe = str(tmp_path / name)
    cmd = ["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror",
           *[os.path.join(ROOT, "examples", f) for f in sources],
           "-I", os.path.join(ROOT, "include"), "-L", LIBDIR, "-lswgl_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_application_compiles_and_links_against_the_library(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_application_renders_k0(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    drawn, color_hash = r.stdout.split()
    assert int(drawn) == 1008 and color_hash == "5e2ecfac3685e7ef"


def test_multi_gpu_c_application_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path, "c4_multi_gpu", ("c4_multi_gpu.c",)))


def _visible_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except (OSError, subprocess.SubprocessError):
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [1, 2, 4, 8])
def test_c_application_renders_c4_on_n_devices(tmp_path, devices):
    """swglSetDeviceCount(N) from plain C: BASELINE config 4 at full size, the colour hash of the frame
    glGetFramePtr returns must be the reference's known-answer value for every N.  Devices beyond the
    visible ones are emulated (members share a GPU): same bands, same sharded uploads, same assembly."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "fullsize_kats.json")) as f:
        want = json.load(f)["C4"]
    env = dict(os.environ)
    if devices > _visible_gpus():
        env["SWGL_B200_GROUP_EMULATE"] = "1"
    r = subprocess.run([_build(tmp_path, "c4_multi_gpu", ("c4_multi_gpu.c",)), str(devices), "708", "3840", "2160", "3"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    n, color_hash, covered, ms = r.stdout.split()
    assert int(n) == devices
    assert color_hash == want["color_fnv"] and int(covered) == want["covered"], r.stdout
