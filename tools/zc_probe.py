"""Experiment: raster write-back stores straight into the pinned host mirror (zero-copy) vs D2H copy."""
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
def frame():
    api.glClear(3)
    api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
def timeit(f, n=20):
    for _ in range(3): f()
    api.swglFinish()
    t0 = time.perf_counter()
    for _ in range(n): f()
    api.swglFinish()
    return (time.perf_counter() - t0) / n * 1e3
def frame_and_read():
    frame(); api.glGetFramePtr()
def frame_finish():
    frame(); api.swglFinish()
print("frame only          %.3f ms" % timeit(frame_finish))
print("frame + D2H copy    %.3f ms" % timeit(frame_and_read))
api.glGetFramePtr.restype = C.c_void_p
hp = api.glGetFramePtr()
ref = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(sc.height, sc.width)).copy()
api.swglSetPeerColorTarget(C.c_uint64(hp))
print("frame, zero-copy    %.3f ms" % timeit(frame_finish))
z = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(sc.height, sc.width))
print("zero-copy image equal:", bool(np.array_equal(z, ref)), api.swglGetLastError())
