"""Shrink a failing sequence of tools/api_fuzz.py: drop one call at a time while the frames still differ.

    python tools/api_fuzz_reduce.py <seed> [lod]

Only calls whose removal cannot make a later call undefined are candidates (draws, viewports, clears, uniforms, texture
calls, reads); programs, vertex arrays and re-specifications stay."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import swgl_b200
from oracle import pyoracle as O
import test_api_fuzz_gpu as F

lod = "lod" in sys.argv
api = swgl_b200.load()
ref = O.Reference(defined_rsqrt=lod)
ops = F.make_ops(int(sys.argv[1]), lod=lod)
msg = F.compare_ops(api, ref, ops, "full", lod=lod)
print(msg or "no mismatch")
if msg:
    removable = {"draw", "points", "viewport", "clear", "clearcolor", "matrix", "tint", "sampler", "wrap", "teximage", "read"}
    changed = True
    while changed:
        changed = False
        for i in range(len(ops) - 1, 0, -1):
            if ops[i][0] not in removable:
                continue
            trial = ops[:i] + ops[i + 1:]
            if F.compare_ops(api, ref, trial, "trial", lod=lod):
                ops = trial
                changed = True
    print(F.compare_ops(api, ref, ops, "reduced", lod=lod))
    print("size", ops[0][3], ops[0][4], "arrays", [(a[0].shape, None if a[1] is None else a[1].shape) for a in ops[0][1]], "textures", [t.shape for t in ops[0][2]])
    for o in ops[1:]:
        if o[0] == "matrix": print(("matrix", np.round(o[1], 3).tolist()))
        elif o[0] in ("teximage",): print((o[0], o[1], o[2].shape))
        elif o[0] == "respecify": print((o[0], o[1], o[2][0].shape, None if o[2][1] is None else o[2][1].shape))
        else: print(o)
