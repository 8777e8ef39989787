"""Does H2D bandwidth depend on which pinned allocation is the source?  (development probe)"""
import torch, time
torch.cuda.set_device(0)
dst = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
bufs = [torch.empty(16 << 20, dtype=torch.uint8).pin_memory() for _ in range(8)]
for b in bufs: b.fill_(1)
s = torch.cuda.Stream()
def bw(b):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        a.record(); dst.copy_(b, non_blocking=True); e.record()
    e.synchronize()
    return b.numel() / a.elapsed_time(e) / 1e6
for rnd in range(4):
    print("round", rnd, " ".join(f"{bw(b):5.1f}" for b in bufs), "GB/s")
    time.sleep(0.2)
h = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
def bw_d2h():
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        a.record(); h.copy_(dst, non_blocking=True); e.record()
    e.synchronize()
    return h.numel() / a.elapsed_time(e) / 1e6
print("d2h", " ".join(f"{bw_d2h():5.1f}" for _ in range(6)))
