"""Quick per-stage timing probe on one GPU (development tool; bench.py is the contract)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S

api = sw.load()
STAGES = ["vertex", "setup_bin", "raster"]


def probe(cfg, reps=10, options=None):
    sc = S.config(cfg) if isinstance(cfg, int) else cfg
    api.glInit(sc.width, sc.height)
    for k, v in (options or {}).items():
        api.swglSetOption(k.encode(), v)
    st = G.setup_scene(api, sc, indexed=sc.indices is not None, init=False)

    def frame():
        api.glClear(3)
        if st["indexed"]:
            api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
        else:
            api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])

    for _ in range(3):
        frame()
    api.swglFinish()
    t0 = time.perf_counter()
    for _ in range(reps):
        frame()
    api.swglFinish()
    dt = (time.perf_counter() - t0) / reps
    api.swglSetOption(b"stage_timing", 1)
    for _ in range(reps):
        frame()
    api.swglFinish()
    nd = api.swglGetOption(b"stage_draws")
    stage = {n: api.swglGetOption(f"stage_ns_{i}".encode()) / 1e3 / max(nd, 1) for i, n in enumerate(STAGES)}
    api.swglSetOption(b"stage_timing", 0)
    s = sw.swglStats()
    api.swglGetStats(C.byref(s))
    err = api.swglGetLastError().decode()
    d = s.as_dict()
    print(f"{sc.name}: frame {dt*1e6:.1f} us  tris {sc.n_triangles} -> {sc.n_triangles/dt/1e6:.1f} Mtri/s  "
          f"shaded {d['shaded']} ({d['shaded']/dt/1e9:.2f} Gfrag/s) tested {d['tested']} pairs {d['tile_pairs']} bands {d['bands']}")
    print("   stages us:", {k: round(v, 1) for k, v in stage.items()}, "sum", round(sum(stage.values()), 1),
          f" roofline frame frac {sc.algorithmic_bytes()/dt/6547.8e9:.4f}", err)
    return dt


if __name__ == "__main__":
    cfgs = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
    for c in cfgs:
        probe(c)
