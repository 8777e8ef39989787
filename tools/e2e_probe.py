"""Where the end-to-end step goes: wall-clock per C-ABI call of the e2e step (host buffers)."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
api = sw.load()
sc = S.config(cfg)
api.glInit(sc.width, sc.height)
for a in sys.argv[2:]:
    k, v = a.split("=")
    api.swglSetOption(k.encode(), int(v))
st = G.setup_scene(api, sc, indexed=True, init=False)
verts = torch.from_numpy(np.ascontiguousarray(sc.vertices)).pin_memory()
idx = torch.from_numpy(np.ascontiguousarray(sc.indices).view(np.int32)).pin_memory()
names = ["respec_v", "respec_i", "clear", "draw", "finish", "getframe"]
acc = {n: 0.0 for n in names}
def step(split):
    t = [time.perf_counter()]
    api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.numel() * 4, C.c_void_p(verts.data_ptr())); t.append(time.perf_counter())
    api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, idx.numel() * 4, C.c_void_p(idx.data_ptr())); t.append(time.perf_counter())
    api.glClear(3); t.append(time.perf_counter())
    api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None); t.append(time.perf_counter())
    if split:
        api.swglFinish()
    t.append(time.perf_counter())
    api.glGetFramePtr(); t.append(time.perf_counter())
    return [b - a for a, b in zip(t, t[1:])]
for split in (1, 0):
    for _ in range(3): step(split)
    N = 20
    tot = np.zeros(6)
    t0 = time.perf_counter()
    for _ in range(N): tot += step(split)
    wall = (time.perf_counter() - t0) / N
    print(f"split={split} wall {wall*1e3:.3f} ms  " + "  ".join(f"{n} {v/N*1e6:.0f}us" for n, v in zip(names, tot)))
print("h2d MB", (verts.numel()+idx.numel())*4/1e6, "d2h MB", sc.width*sc.height*4/1e6)
