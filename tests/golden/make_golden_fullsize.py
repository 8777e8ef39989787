"""Generate tests/golden/fullsize_kats.json: BASELINE configs 2..5 at full size through the COMPILED,
UNMODIFIED reference (minutes of CPU time; run in the authoring container, needs /root/reference):

    python tests/golden/make_golden_fullsize.py

The GPU tests compare the CUDA path's images with these hashes, so full-size parity does not need the
reference on the GPU box.  Fragment counters come from the restatement after its image has been
checked equal to the reference's."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402
from swgl_b200 import scenes as S  # noqa: E402


def main():
    ref, rest = O.Reference(), O.Restatement()
    out = {}
    for n in (2, 3, 4, 5):
        scene = S.config(n)
        t0 = time.time()
        col, dep = ref.render(scene)
        t1 = time.time()
        rc, rd, stats = rest.render(scene)
        cmp = O.compare(col, dep, rc, rd)
        assert cmp["depth_mismatch"] == 0 and cmp["color_mismatch"] == 0, (n, cmp)
        out[f"C{n}"] = {
            "scene": scene.name, "width": scene.width, "height": scene.height,
            "color_fnv": f"{rest.fnv(col):016x}", "depth_fnv": f"{rest.fnv(dep):016x}",
            "covered": int((dep.view(np.uint32) != 0).sum()), "nan": int(np.isnan(dep).sum()),
            "tested": stats["tested"], "shaded": stats["shaded"],
            "reference_seconds": round(t1 - t0, 2),
        }
        print(n, out[f"C{n}"], flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_kats.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
