"""Per-call wall clock of the pipelined e2e step (swglFrameSubmit / swglFrameWait, two geometry sets)."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
api = sw.load()
sc = S.config(cfg)
api.glInit(sc.width, sc.height)
st = G.setup_scene(api, sc, indexed=True, init=False)
sets = [G.add_geometry(api, sc, indexed=True, named=True)[:3] for _ in range(2)]
verts = torch.from_numpy(np.ascontiguousarray(sc.vertices)).pin_memory()
idx = torch.from_numpy(np.ascontiguousarray(sc.indices).view(np.int32)).pin_memory()
names = ["bind", "respec_v", "respec_i", "clear+draw", "submit", "wait_prev"]
def step(i, prev):
    t = [time.perf_counter()]
    vao, vbo, ebo = sets[i & 1]
    api.glBindVertexArray(vao); api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo); api.glBindBuffer(G.GL_ELEMENT_ARRAY_BUFFER, ebo); t.append(time.perf_counter())
    api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.numel() * 4, C.c_void_p(verts.data_ptr())); t.append(time.perf_counter())
    api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, idx.numel() * 4, C.c_void_p(idx.data_ptr())); t.append(time.perf_counter())
    api.glClear(3); api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None); t.append(time.perf_counter())
    tk = api.swglFrameSubmit(); t.append(time.perf_counter())
    if prev: api.swglFrameWait(prev)
    t.append(time.perf_counter())
    return tk, [b - a for a, b in zip(t, t[1:])]
prev = 0
for i in range(4): prev, _ = step(i, prev)
N = 20
tot = np.zeros(6)
t0 = time.perf_counter()
for i in range(N):
    prev, d = step(i, prev); tot += d
api.swglFrameWait(prev)
wall = (time.perf_counter() - t0) / N
print(f"wall {wall*1e3:.3f} ms  " + "  ".join(f"{n} {v/N*1e6:.0f}us" for n, v in zip(names, tot)), "wt_draws", api.swglGetOption(b"wt_draws"), api.swglGetLastError())
