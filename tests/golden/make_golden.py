"""Generate tests/golden/oracle_kats.json from the COMPILED, UNMODIFIED reference.

Run in the authoring container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
Fragment counters come from the restatement (the reference does not count), after the
restatement's image has been checked equal to the reference's.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as O  # noqa: E402
from test_oracle import golden_scenes  # noqa: E402


def main():
    ref, rest = O.Reference(), O.Restatement()
    out = {}
    for name, scene in golden_scenes().items():
        col, dep = ref.render(scene)
        rc, rd, stats = rest.render(scene)
        cmp = O.compare(col, dep, rc, rd)
        assert cmp["depth_mismatch"] == 0 and cmp["color_mismatch"] == 0, (name, cmp)
        out[name] = {
            "scene": scene.name, "width": scene.width, "height": scene.height,
            "color_fnv": f"{rest.fnv(col):016x}", "depth_fnv": f"{rest.fnv(dep):016x}",
            "covered": int((dep.view(np.uint32) != 0).sum()), "nan": int(np.isnan(dep).sum()),
            "tested": stats["tested"], "shaded": stats["shaded"], "prims_out": stats["prims_out"],
        }
        print(name, out[name])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_kats.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
