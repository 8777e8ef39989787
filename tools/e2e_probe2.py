"""Which part of bench.py's sequence disturbs the end-to-end step: python tools/e2e_probe2.py [flush] [stage] [sampler] [events]"""
import sys, time, ctypes as C, subprocess
sys.path.insert(0, ".")
import numpy as np, torch
import swgl_b200 as sw
from swgl_b200 import gl as G, scenes as S
flags = set(sys.argv[1:])
torch.cuda.set_device(0)
api = sw.load()
sc = S.config(4)
api.swglSetDevice(0)
api.glInit(sc.width, sc.height)
st = G.setup_scene(api, sc, indexed=True, init=False)
if "pinfirst" in flags:
    verts = torch.from_numpy(np.ascontiguousarray(sc.vertices)).pin_memory()
    idx = torch.from_numpy(np.ascontiguousarray(sc.indices).view(np.int32)).pin_memory()
stream = torch.cuda.ExternalStream(api.swglGetStream(), device=torch.device("cuda", 0))
def frame():
    api.glClear(3); api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
proc = None
if "sampler" in flags:
    proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm", "--format=csv,noheader", "-lms", "20"], stdout=subprocess.DEVNULL)
for _ in range(3): frame()
api.swglFinish()
if "flush" in flags:
    flush = torch.empty((256 << 20) // 4, dtype=torch.int32, device="cuda:0")
    for _ in range(20):
        with torch.cuda.stream(stream):
            flush.zero_()
        frame()
    api.swglFinish(); torch.cuda.synchronize()
if "events" in flags:
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in evs:
        a.record(stream); frame(); b.record(stream)
    api.swglFinish(); torch.cuda.synchronize()
if "stage" in flags:
    api.swglSetOption(b"stage_timing", 1)
    for _ in range(20): frame()
    api.swglFinish()
    api.swglSetOption(b"stage_timing", 0)
if "pinfirst" not in flags:
    verts = torch.from_numpy(np.ascontiguousarray(sc.vertices)).pin_memory()
    idx = torch.from_numpy(np.ascontiguousarray(sc.indices).view(np.int32)).pin_memory()
if "wc" in flags or "pin" in flags:
    def stage(a):
        a = np.ascontiguousarray(a)
        ptr = api.swglHostAlloc(a.nbytes, 1 if "wc" in flags else 0)
        C.memmove(ptr, a.ctypes.data, a.nbytes)
        class V:  # minimal stand-in for the torch tensor interface used below
            def numel(self): return a.nbytes // 4
            def data_ptr(self): return ptr
        return V()
    verts = stage(sc.vertices); idx = stage(sc.indices)
names = ["respec_v", "respec_i", "clear", "draw", "getframe"]
def step():
    t = [time.perf_counter()]
    api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.numel() * 4, C.c_void_p(verts.data_ptr())); t.append(time.perf_counter())
    api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, idx.numel() * 4, C.c_void_p(idx.data_ptr())); t.append(time.perf_counter())
    api.glClear(3); t.append(time.perf_counter())
    api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None); t.append(time.perf_counter())
    api.glGetFramePtr(); t.append(time.perf_counter())
    return [b - a for a, b in zip(t, t[1:])]
for _ in range(2): step()
tot = np.zeros(5); t0 = time.perf_counter(); per = []
for _ in range(int(__import__("os").environ.get("STEPS", "20"))):
    r = step(); tot += r; per.append(round((r[0] + r[1]) * 1e6))
print("respec us per step:", per)
wall = (time.perf_counter() - t0) / 20
print(sorted(flags), f"wall {wall*1e3:.3f} ms  " + "  ".join(f"{n} {v/20*1e6:.0f}us" for n, v in zip(names, tot)), "wt_draws", api.swglGetOption(b"mirror_synced"))
if proc: proc.terminate()
