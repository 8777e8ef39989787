"""Generate tests/golden/lod_kats.json: frames of mip-mapped scenes from the reference COMPILED WITH THE DEFINED rsqrt
(oracle/ref_shim.c -DSWGLREF_DEFINED_RSQRT: `long` is 32 bits for the span of the #include, no other change) --
the variant in which the level of detail is not undefined behaviour.

Run in the authoring container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden_lod.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as O  # noqa: E402
from test_oracle import lod_scenes  # noqa: E402


def main():
    ref, rest = O.Reference(defined_rsqrt=True), O.Restatement()
    out = {}
    for name, scene in lod_scenes().items():
        col, dep = ref.render(scene, mipmaps=True)
        rc, rd, stats = rest.render(scene, mipmaps=True)
        cmp = O.compare(col, dep, rc, rd)
        assert cmp["depth_mismatch"] == 0 and cmp["color_mismatch"] == 0, (name, cmp)
        base, _ = ref.render(scene)
        out[name] = {
            "scene": scene.name, "width": scene.width, "height": scene.height,
            "color_fnv": f"{rest.fnv(col):016x}", "depth_fnv": f"{rest.fnv(dep):016x}",
            "covered": int((dep.view(np.uint32) != 0).sum()), "differs_from_base_level": int((col != base).sum()),
            "tested": stats["tested"], "shaded": stats["shaded"],
        }
        print(name, out[name])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lod_kats.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
