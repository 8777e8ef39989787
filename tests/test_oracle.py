"""CPU: pin the C restatement (oracle/swgl_oracle.c) against the compiled, unmodified reference
and against golden hashes (tests/golden/oracle_kats.json, produced by tests/golden/make_golden.py
from the compiled reference; K0/K1/K2 also equal SURVEY.md appendix C)."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from swgl_b200 import scenes as S

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_kats.json")


def golden_scenes():
    sc = {
        "k0_single": S.single_triangle(),
        "k1_config1": S.random_triangles(),
        "k2_config1_tex": S.random_triangles(textured=True),
        "near_clip_alpha": S.random_triangles(2000, near_cross=True, alpha=None, centre_range=1.3, seed=7),
        "grid32": S.grid_mesh(32, 320, 200),
        "grid32_tex": S.grid_mesh(32, 320, 200, textured=True),
        "grid40_matrix_a05": S.grid_mesh(40, 333, 211, alpha=0.5, use_matrix=True),
    }
    vp = S.random_triangles(1500, near_cross=True, alpha=None, centre_range=1.3, seed=99)
    vp.viewport = (37, 11, 301, 257)
    sc["viewport_37_11"] = vp
    return sc


with open(GOLDEN) as f:
    KATS = json.load(f)


def lod_scenes():
    """Mip-mapped scenes of the DEFINED level-of-detail variant (tests/golden/make_golden_lod.py): triangles from far
    larger than 40 pixels (level below 1) to a few pixels (level beyond the chain), both wrap modes, near clipping."""
    sc = {
        "lod_grid9": S.grid_mesh(9, 320, 240, textured=True),
        "lod_grid24": S.grid_mesh(24, 320, 240, textured=True),
        "lod_grid60_clamp": S.grid_mesh(60, 320, 240, textured=True),
        "lod_random_near": S.random_triangles(600, 320, 240, seed=5, extent=0.12, textured=True, near_cross=True),
    }
    sc["lod_grid60_clamp"].tex_wrap = "clamp"
    return sc


LOD_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "lod_kats.json")
LOD_KATS = json.load(open(LOD_GOLDEN)) if os.path.exists(LOD_GOLDEN) else {}


@pytest.mark.parametrize("name", sorted(KATS))
def test_restatement_matches_golden(restatement, name):
    scene = golden_scenes()[name]
    col, dep, stats = restatement.render(scene)
    k = KATS[name]
    assert f"{restatement.fnv(col):016x}" == k["color_fnv"]
    assert f"{restatement.fnv(dep):016x}" == k["depth_fnv"]
    assert int((dep.view(np.uint32) != 0).sum()) == k["covered"]
    assert stats["tested"] == k["tested"] and stats["shaded"] == k["shaded"]


@pytest.mark.parametrize("name", ["k0_single", "k1_config1", "near_clip_alpha", "grid32_tex", "viewport_37_11"])
def test_restatement_matches_compiled_reference(restatement, reference, name):
    scene = golden_scenes()[name]
    c1, d1, _ = restatement.render(scene)
    c2, d2 = reference.render(scene)
    cmp = O.compare(c1, d1, c2, d2)
    assert cmp["depth_mismatch"] == 0 and cmp["coverage_mismatch"] == 0 and cmp["color_mismatch"] == 0, cmp


@pytest.mark.parametrize("name", sorted(LOD_KATS))
def test_restatement_mip_levels_match_golden(restatement, name):
    """glGenerateMipmap + per-triangle level of detail in the restatement (swgl_oracle.c: swglo_build_mipmaps, mip_level,
    sample_lod) against the frames of the reference built with the defined rsqrt."""
    scene = lod_scenes()[name]
    col, dep, stats = restatement.render(scene, mipmaps=True)
    k = LOD_KATS[name]
    assert f"{restatement.fnv(col):016x}" == k["color_fnv"] and f"{restatement.fnv(dep):016x}" == k["depth_fnv"]
    assert stats["tested"] == k["tested"] and stats["shaded"] == k["shaded"]
    base, _, _ = restatement.render(scene)
    assert int((col != base).sum()) == k["differs_from_base_level"] > 10000


@pytest.mark.parametrize("name", sorted(LOD_KATS))
def test_restatement_mip_levels_match_the_defined_rsqrt_reference(restatement, name):
    try:
        ref = O.Reference(defined_rsqrt=True)
    except Exception as e:                                  # pragma: no cover
        pytest.skip(f"oracle/_ref/libswgl_ref_lod.so not available: {e}")
    scene = lod_scenes()[name]
    c1, d1, _ = restatement.render(scene, mipmaps=True)
    c2, d2 = ref.render(scene, mipmaps=True)
    cmp = O.compare(c1, d1, c2, d2)
    assert cmp["depth_mismatch"] == 0 and cmp["coverage_mismatch"] == 0 and cmp["color_mismatch"] == 0, cmp


def _point_scenes():
    sc = {
        "colour_near": (S.random_triangles(900, 320, 240, seed=3, alpha=None, near_cross=True, centre_range=1.3), None),
        "colour_viewport": (S.random_triangles(900, 320, 240, seed=4, alpha=None), (13, 7, 150, 90)),
        "textured": (S.random_triangles(900, 320, 240, seed=6, textured=True), None),
        "matrix": (S.grid_mesh(30, 333, 211, alpha=0.5, use_matrix=True), None),
    }
    for scene, vp in sc.values():
        scene.viewport = vp
        if scene.indices is not None:
            scene.vertices, scene.indices = scene.deindexed(), None
    return {k: v[0] for k, v in sc.items()}


@pytest.mark.parametrize("name", ["colour_near", "colour_viewport", "textured", "matrix"])
def test_restatement_points_match_compiled_reference(restatement, reference, name):
    """GL_POINTS (swgl.c:3496-3608; swglo_draw_points): 100 triangles first -- the reference only allocates a program's
    fragment inputs when a triangle is shaded -- then every other vertex of the stream as a point."""
    scene = _point_scenes()[name]
    kw = dict(count=300, points=(300, len(scene.vertices) - 300))
    c1, d1, _ = restatement.render(scene, **kw)
    c2, d2 = reference.render(scene, **kw)
    cmp = O.compare(c1, d1, c2, d2)
    assert cmp["depth_mismatch"] == 0 and cmp["coverage_mismatch"] == 0 and cmp["color_mismatch"] == 0, cmp
    t1, _, _ = restatement.render(scene, count=300)
    assert int((t1 != c1).sum()) > 300            # the points are really there


def test_restatement_points_sample_with_the_last_triangles_level(restatement):
    """Under the defined level of detail a point samples with the level the last triangle left (a global, swgl.c:3314)."""
    try:
        ref = O.Reference(defined_rsqrt=True)
    except Exception as e:                                  # pragma: no cover
        pytest.skip(f"oracle/_ref/libswgl_ref_lod.so not available: {e}")
    scene = _point_scenes()["textured"]
    kw = dict(count=300, points=(300, len(scene.vertices) - 300), mipmaps=True)
    c1, d1, _ = restatement.render(scene, **kw)
    c2, d2 = ref.render(scene, **kw)
    cmp = O.compare(c1, d1, c2, d2)
    assert cmp["depth_mismatch"] == 0 and cmp["coverage_mismatch"] == 0 and cmp["color_mismatch"] == 0, cmp
    base, _, _ = restatement.render(scene, count=300, points=(300, len(scene.vertices) - 300))
    assert int((base != c1).sum()) > 10000


def test_survey_kats():
    """SURVEY.md appendix C values, measured independently during the survey."""
    assert KATS["k0_single"]["color_fnv"] == "5e2ecfac3685e7ef" and KATS["k0_single"]["covered"] == 1008
    assert KATS["k1_config1"]["color_fnv"] == "c448a459df88d572"
    assert KATS["k1_config1"]["depth_fnv"] == "00fd259510614c5e"
    assert KATS["k1_config1"]["covered"] == 304414
    assert KATS["k1_config1"]["tested"] == 1878617 and KATS["k1_config1"]["shaded"] == 723510
    assert KATS["k2_config1_tex"]["color_fnv"] == "ebe51dc3610b6c83"


def test_no_clear_and_partial_draw(restatement, reference):
    scene = S.random_triangles(300, 320, 240, seed=5, alpha=0.5)
    c1, d1, _ = restatement.render(scene, clear=False, fill=(0x80402010, 0.0), first=30, count=600)
    c2, d2 = reference.render(scene, clear=False, fill=(0x80402010, 0.0), first=30, count=600)
    cmp = O.compare(c1, d1, c2, d2)
    assert cmp["depth_mismatch"] == 0 and cmp["color_mismatch"] == 0, cmp


@pytest.mark.parametrize("viewport", [(0, -40, 200, 152), (0, 30, 200, 152), (-10, -25, 220, 212), (13, -7, 150, 300),
                                      (89, -24, 101, 10), (0, 160, 200, 50)])
def test_restatement_folds_rows_like_the_compiled_reference(restatement, reference, viewport):
    """Viewports that leave the framebuffer rows (swgl.c:3386: rows outside land on row Height-1; a viewport that ends
    below row 0 also loses its row limit, 3344 / 3356).  The restatement's fragment counts for such draws are what
    tests/test_api_edge_gpu.py holds the library's folded draws to."""
    scene = S.random_triangles(400, 200, 152, seed=21, near_cross=True, alpha=None, centre_range=1.2)
    scene.viewport = viewport
    c1, d1, stats = restatement.render(scene)
    c2, d2 = reference.render(scene)
    cmp = O.compare(c1, d1, c2, d2)
    assert cmp["depth_mismatch"] == 0 and cmp["coverage_mismatch"] == 0 and cmp["color_mismatch"] == 0, cmp
    assert stats["tested"] > 0 and int((d2[151].view(np.uint32) != 0).sum()) > 0


def test_restatement_matches_fullsize_golden_c2(restatement):
    """BASELINE config 2 at full size (1080p, 100,352 triangles): the restatement against the hashes the
    compiled reference produced (tests/golden/fullsize_kats.json); also the survey's K3 values."""
    with open(os.path.join(os.path.dirname(GOLDEN), "fullsize_kats.json")) as f:
        full = json.load(f)
    k = full["C2"]
    col, dep, stats = restatement.render(S.config(2))
    assert f"{restatement.fnv(col):016x}" == k["color_fnv"] == "9254eae11a6abc99"
    assert f"{restatement.fnv(dep):016x}" == k["depth_fnv"]
    assert int((dep.view(np.uint32) != 0).sum()) == k["covered"] == 1538557
    assert stats["tested"] == k["tested"] and stats["shaded"] == k["shaded"]
    assert full["C4"]["color_fnv"] == "56d0d4e1a9cd44f1" and full["C4"]["covered"] == 5202381   # K4
    assert full["C5"]["covered"] == 20795225                                                    # K5


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU path, oracle/_ref): one JSON line with the
    keys the measurement contract names; C1 so that it takes a second."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "triangles/s" and d["unit"] == "triangles/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("C1")
