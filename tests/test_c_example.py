"""The drop-in claim, literally: a plain C application (two translation units that both include
swgl.h) compiled with gcc and linked against libswgl_b200.so instead of swgl.c.
CPU: it compiles and links.  GPU: it renders known-answer scene K0 with the reference's result."""
import os
import subprocess

import pytest

from swgl_b200._lib import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.dirname(LIB_PATH)


def _build(tmp_path):
    exe = str(tmp_path / "k0")
    cmd = ["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Werror",
           os.path.join(ROOT, "examples", "k0_triangle.c"), os.path.join(ROOT, "examples", "k0_hash.c"),
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
