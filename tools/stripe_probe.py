"""Per-stage timing of one rank's share of a sort-first frame, emulated on one GPU:
    python tools/stripe_probe.py <cfg> <world> [rank] [band_rows]"""
import sys
sys.path.insert(0, ".")
import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S

cfg = int(sys.argv[1]); world = int(sys.argv[2])
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
band = int(sys.argv[4]) if len(sys.argv) > 4 else 1
api = sw.load()
sc = S.config(cfg)
api.glInit(sc.width, sc.height)
st = G.setup_scene(api, sc, indexed=sc.indices is not None, init=False)
for a in sys.argv[5:]:
    k, v = a.split("=")
    api.swglSetOption(k.encode(), int(v))
api.swglSetStripe(rank, world, band)
def frame():
    api.glClear(3)
    if st["indexed"]:
        api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
    else:
        api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
for _ in range(3):
    frame()
api.swglFinish()
api.swglSetOption(b"stage_timing", 1)
for _ in range(20):
    frame()
api.swglFinish()
nd = api.swglGetOption(b"stage_draws")
print(f"C{cfg} rank {rank}/{world}:", {n: round(api.swglGetOption(f"stage_ns_{i}".encode()) / 1e3 / max(nd, 1), 1) for i, n in enumerate(["vertex", "setup_bin", "raster"])}, api.swglGetLastError())
