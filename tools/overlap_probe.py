"""What slows an upload down when it overlaps other work? (development probe)"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S

api = sw.load()
sc = S.config(4)
api.glInit(sc.width, sc.height)
api.swglSetOption(b"host_mirror", 0)
st = G.setup_scene(api, sc, indexed=True, init=False)
sets = [G.add_geometry(api, sc, indexed=True, named=True)[:3] for _ in range(2)]
verts = torch.from_numpy(np.ascontiguousarray(sc.vertices)).pin_memory()
def bind(i):
    vao, vbo, ebo = sets[i]
    api.glBindVertexArray(vao); api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo); api.glBindBuffer(G.GL_ELEMENT_ARRAY_BUFFER, ebo)
def frame():
    api.glClear(3); api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
def respec():
    t0 = time.perf_counter()
    api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.numel() * 4, C.c_void_p(verts.data_ptr()))
    return (time.perf_counter() - t0) * 1e6
bind(0); frame(); api.swglFinish()
bind(1); print("A alone            %.0f us" % respec())
bind(1); print("A alone            %.0f us" % respec())
# B: set 0 drawn 10 times (queued), upload into set 1 meanwhile
bind(0)
for _ in range(10): frame()
bind(1); tb = respec(); api.swglFinish()
print("B under 10 frames   %.0f us" % tb)
# C: D2H DMA of 33 MB x 8 on a torch stream, upload meanwhile
dev = torch.empty(sc.width * sc.height, dtype=torch.int32, device="cuda")
host = torch.empty(sc.width * sc.height, dtype=torch.int32).pin_memory()
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(s2):
    for _ in range(8): host.copy_(dev, non_blocking=True)
tc = respec(); torch.cuda.synchronize()
print("C under D2H DMA     %.0f us" % tc)
# D: a long pure-compute torch kernel meanwhile
x = torch.randn(8192, 8192, device="cuda")
torch.cuda.synchronize()
with torch.cuda.stream(s2):
    for _ in range(4): y = x @ x
td = respec(); torch.cuda.synchronize()
print("D under GEMMs       %.0f us" % td)
