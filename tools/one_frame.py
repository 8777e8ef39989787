"""Render a few frames of one BASELINE config (profiling target for ncu)."""
import sys
sys.path.insert(0, ".")
import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
api = sw.load()
sc = S.config(cfg)
api.glInit(sc.width, sc.height)
for a in sys.argv[3:]:
    k, v = a.split("=")
    api.swglSetOption(k.encode(), int(v))
st = G.setup_scene(api, sc, indexed=sc.indices is not None, init=False)
for _ in range(frames):
    api.glClear(3)
    if st["indexed"]:
        api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
    else:
        api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])
api.swglFinish()
print("ok", api.swglGetLastError())
