#!/usr/bin/env python
"""bench.py -- the measurement contract for the swgl draw-call path on B200.

    python bench.py --gpus 1 --steps K --warmup W                 # CUDA path (this repo)
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # the reference's CPU path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A *step* is one frame of the workload: glClear(COLOR|DEPTH) + one indexed draw of BASELINE
config 4 (4K, 1,002,528 small depth-tested triangles) unless --config says otherwise.

Own arm, one JSON line on rank 0:
  value        triangles/s with every input resident in HBM; per-step CUDA events on the
               library's stream, L2 flushed (256 MiB write) between steps, max over ranks
  e2e          the same frame through the C ABI with HOST buffers: pinned vertex/index arrays
               re-specified every step (H2D), clear + draw, glGetFramePtr (D2H of the image)
  roofline     dominant kernel (k_raster): algorithmic bytes per launch (W*H*8: colour + depth
               written once with the clear fused) / its CUDA-event duration,
               against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the unmodified reference (oracle/_ref) on one host core, bounded sample
N > 1: sort-first by tile-row bands; geometry replicated; every rank stores its finished tiles
straight into rank 0's colour buffer over NVLink (CUDA IPC peer mapping), torch.distributed
(NCCL) provides the barriers and the max-over-ranks reduction.  Same frame at every N
(strong scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "triangles/s"
STAGES = ["k_vertex", "k_setup_bin", "k_raster"]
L2_FLUSH_BYTES = 256 << 20


def band_rows_for(height: int, world: int) -> int:
    """Sort-first ownership: >= 16 interleaved bands of 32 rows per rank."""
    if world <= 1:
        return 1
    return max(1, ((height + 31) // 32) // (world * 16))


def config_dict(cfg: int, scene, world: int) -> dict:
    """`config` of the JSON line: the same dict in both arms (the driver compares them)."""
    from swgl_b200 import scenes as S

    br = band_rows_for(scene.height, world)
    return {"workload": S.CONFIG_NAMES[cfg], "triangles": scene.n_triangles,
            "framebuffer": f"{scene.width}x{scene.height} RGBA8 + f32 depth",
            "l2": "GPU arm: flushed between steps (256 MiB write, untimed); CPU arm: n/a",
            "parallelism": f"GPU arm: sort-first tile-row bands x{world}" + (f", band={br} tile rows, peer stores to rank 0" if world > 1 else "")
                           + "; CPU arm: one host core"}


def golden_for(cfg: int):
    """Known-answer frame of a BASELINE config rendered by the compiled, unmodified reference
    (tests/golden/*.json, committed with the scripts that made them)."""
    name, key = ("oracle_kats.json", "k1_config1") if cfg == 1 else ("fullsize_kats.json", f"C{cfg}")
    path = os.path.join(ROOT, "tests", "golden", name)
    try:
        with open(path) as f:
            return json.load(f).get(key)
    except OSError:
        return None


def ncu_counters(cfg: int, kernel: str):
    """Per-launch counters of one `ncu --set full` capture (profiles/ncu_counters.json, written by
    tools/ncu_summary.py with the commit it was captured at); None when there is no capture of this config."""
    path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except OSError:
        return None, None
    return d.get(f"C{cfg}", {}).get(kernel), d.get("captured_at")


def parity_block(api, scene, cfg: int, stats: dict) -> dict:
    """The frame just rendered against the reference's known-answer frame of this config: FNV-1a64 of the colour
    words and of the depth bits, covered pixels, tested / shaded fragments."""
    want = golden_for(cfg)
    W, H = scene.width, scene.height
    n = W * H
    col = api.glGetFramePtr()
    color_fnv = api.swglHashWords(col, n)
    dep = api.swglGetDepthPtr()
    depth_fnv = api.swglHashWords(dep, n)
    d = np.ctypeslib.as_array(dep, shape=(H, W))
    covered = int((d.view(np.uint32) != 0).sum())
    out = {"color_fnv": f"{color_fnv:016x}", "depth_fnv": f"{depth_fnv:016x}", "covered": covered,
           "tested": int(stats["tested"]), "shaded": int(stats["shaded"])}
    if want is None:
        out["golden"] = None
        return out
    out["golden"] = "tests/golden (compiled reference)"
    out["color_equal"] = out["color_fnv"] == want["color_fnv"]
    out["depth_equal"] = out["depth_fnv"] == want["depth_fnv"]
    out["covered_delta"] = covered - int(want["covered"])
    out["tested_delta"] = out["tested"] - int(want["tested"])
    out["shaded_delta"] = out["shaded"] - int(want["shaded"])
    out["bit_exact"] = bool(out["color_equal"] and out["depth_equal"] and out["covered_delta"] == 0
                            and out["tested_delta"] == 0 and out["shaded_delta"] == 0)
    return out


def measure_config(api, sw, G, S, torch, cfg: int, steps: int, stream_of, flush, peak: float) -> dict:
    """One of the other BASELINE configs on this GPU: device-resident frame time (CUDA events, L2 flushed),
    per-stage times, roofline fraction of the frame's algorithmic bytes, parity against the golden frame."""
    scene = S.config(cfg)
    api.glInit(scene.width, scene.height)
    err = api.swglGetLastError().decode()
    if err:
        return {"error": err}
    st = G.setup_scene(api, scene, indexed=scene.indices is not None, init=False)
    stream = stream_of()

    def frame():
        api.glClear(G.GL_COLOR_BUFFER_BIT | G.GL_DEPTH_BUFFER_BIT)
        if st["indexed"]:
            api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
        else:
            api.glDrawArrays(G.GL_TRIANGLES, 0, st["n_draw"])

    for _ in range(3):
        frame()
    api.swglFinish()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        with torch.cuda.stream(stream):
            flush.zero_()
        a.record(stream)
        frame()
        b.record(stream)
    api.swglFinish()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    api.swglSetOption(b"stage_timing", 1)
    for _ in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        frame()
    api.swglFinish()
    nd = max(1, api.swglGetOption(b"stage_draws"))
    stage_us = {n: api.swglGetOption(f"stage_ns_{i}".encode()) / 1e3 / nd for i, n in enumerate(STAGES)}
    api.swglSetOption(b"stage_timing", 0)
    frame()
    stats = sw.swglStats()
    api.swglGetStats(C.byref(stats))
    sd = stats.as_dict()
    par = parity_block(api, scene, cfg, sd)
    err = api.swglGetLastError().decode()
    out = {"workload": S.CONFIG_NAMES[cfg], "triangles": scene.n_triangles, "frame_ms": ms,
           "triangles_per_s": scene.n_triangles / (ms * 1e-3), "shaded_fragments_per_s": sd["shaded"] / (ms * 1e-3),
           "tested_fragments_per_s": sd["tested"] / (ms * 1e-3), "stage_us": stage_us,
           "frame_algorithmic_bytes": scene.algorithmic_bytes(),
           "frame_frac": scene.algorithmic_bytes() / (ms * 1e-3) / 1e9 / peak,
           "raster_frac": scene.width * scene.height * 8 / (stage_us["k_raster"] * 1e-6) / 1e9 / peak if stage_us["k_raster"] > 0 else None,
           "parity": par}
    if err:
        out["error"] = err
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        with open(self.tmp.name) as f:
            rows = [r.strip().split(",") for r in f if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, n in enumerate(names):
                if len(r) > 5 + k and r[5 + k].strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # NCCL's banner ("NCCL version ...") goes to stdout at the VERSION level: keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return dist, rank, world, local


def e2e_record(n_tris, e2e_s, verts_bytes, idx_bytes, scene):
    return {"value": n_tris / e2e_s, "unit": METRIC, "ms_per_step": e2e_s * 1e3,
            "h2d_bytes_per_step": int(verts_bytes + idx_bytes),
            "host_buffers": "page-locked, write-combined (swglHostAlloc)",
            "d2h_bytes_per_step": scene.width * scene.height * 4}


def group_e2e(api, sw, G, scene, args, world, local):
    """Rank 0, N > 1: the end-to-end step through the library's device group (swglSetDeviceCount)."""
    from swgl_b200 import multigpu

    api.swglSetDevice(0)
    api.swglSetDeviceCount(world)
    api.glInit(scene.width, scene.height)
    err = api.swglGetLastError().decode()
    if err or api.swglGetOption(b"device_count") != world:
        raise RuntimeError(f"device group of {world} could not be created: {err}")
    st = G.setup_scene(api, scene, indexed=True, init=False)
    verts = multigpu.HostArray(api, scene.vertices)
    idx = multigpu.HostArray(api, scene.indices)

    def step():
        api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.nbytes, C.c_void_p(verts.ptr))
        api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, idx.nbytes, C.c_void_p(idx.ptr))
        api.glClear(G.GL_COLOR_BUFFER_BIT | G.GL_DEPTH_BUFFER_BIT)
        api.glDrawElements(G.GL_TRIANGLES, st["n_draw"], G.GL_UNSIGNED_INT, None)
        api.glGetFramePtr()

    for _ in range(max(args.warmup, 3)):
        step()
    api.swglFinish()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    e2e_s = (time.perf_counter() - t0) / args.steps
    out = e2e_record(scene.n_triangles, e2e_s, verts.nbytes, idx.nbytes, scene)
    out["driver"] = (f"rank 0 drives all {world} GPUs from one host thread: swglSetDeviceCount({world}) + the reference's calls; "
                     "uploads cross PCIe once (1/N per device) and are completed over NVLink, every device writes its bands "
                     "of the frame into one pinned host mirror")
    stats = sw.swglStats()
    api.swglGetStats(C.byref(stats))
    out["parity"] = parity_block(api, scene, args.config, stats.as_dict())
    out["kernel_launches_per_step"] = None
    verts.free()
    idx.free()
    err = api.swglGetLastError().decode()
    if err:
        out["error"] = err
    return out


def ranks_e2e(api, G, multigpu, dist, torch, scene, args, rank, world, local, frame, barrier, assembled_equals_single, peer):
    """The end-to-end step with one process per GPU (N = 1, or --e2e-mode ranks): at N > 1 every rank uploads 1/N
    of the arrays over its own PCIe link, an NCCL all-gather over NVLink on the library's stream replicates them,
    and every rank writes its bands of the frame into one shared host segment."""
    n_tris = scene.n_triangles
    shared = None
    if world > 1:
        if peer is not None:
            peer.close()
        try:
            shared = multigpu.SharedFrameMirror(api, dist, rank, world, scene.width, scene.height)
        except (RuntimeError, OSError):
            shared = None
            peer = multigpu.PeerColorTarget(api, dist, rank, world)
    verts = multigpu.HostArray(api, scene.vertices)
    idx = multigpu.HostArray(api, scene.indices)
    sharded = None
    if world > 1 and multigpu.ShardedUpload.divisible(verts.nbytes, world) and multigpu.ShardedUpload.divisible(idx.nbytes, world):
        sharded = multigpu.ShardedUpload(api, dist, rank, world, torch.device("cuda", local))

    def e2e_step():
        nonlocal sharded
        if sharded is not None:
            try:
                sharded.upload(G.GL_ARRAY_BUFFER, verts)
                sharded.upload(G.GL_ELEMENT_ARRAY_BUFFER, idx)
            except Exception as exc:      # the same call fails the same way on every rank: all fall back together
                print(f"bench: sharded upload unavailable ({exc!r}); every rank uploads the whole arrays", file=sys.stderr)
                sharded = None
        if sharded is None:
            api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.nbytes, C.c_void_p(verts.ptr))
            api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, idx.nbytes, C.c_void_p(idx.ptr))
        frame()
        if dist is not None:
            api.swglFinish()
            dist.barrier()
        if rank == 0:
            api.glGetFramePtr()   # N = 1: wait for the written-through mirror; N > 1: the shared segment (or sync + D2H)

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if dist is not None:
        t = torch.tensor([e2e_s], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = e2e_record(n_tris, e2e_s, verts.nbytes * (1 if sharded is not None or world == 1 else world),
                     idx.nbytes * (1 if sharded is not None or world == 1 else world), scene)
    if world > 1:
        e2e["driver"] = "one process per GPU (torch.distributed)"
        e2e["upload"] = ("1/N of the arrays per rank over its own PCIe link + NCCL all-gather over NVLink" if sharded is not None
                         else "every rank uploads the whole arrays")
        if sharded is not None:
            e2e["nvlink_bytes_per_step_per_rank"] = int(verts.nbytes + idx.nbytes) * (world - 1) // world
        e2e["assembly"] = ("shared host frame mirror: every rank writes its bands over its own PCIe link" if shared is not None
                           else "peer stores into rank 0's HBM, then one D2H copy on rank 0")
        if shared is not None:
            e2e["host_mirror_check"] = assembled_equals_single()
            shared.close()
    verts.free()
    idx.free()
    if peer is not None:
        peer.close()
    return e2e


# ---------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import swgl_b200 as sw
    from swgl_b200 import gl as G, scenes as S
    from swgl_b200 import multigpu

    dist, rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    api = sw.load()
    scene = S.config(args.config)
    api.swglSetDevice(local)
    api.glInit(scene.width, scene.height)
    err = api.swglGetLastError().decode()
    if err:
        raise RuntimeError(f"swgl_b200: {err}")  # no CPU fallback: fail loudly
    st = G.setup_scene(api, scene, indexed=True, init=False)
    n_draw = st["n_draw"]
    n_tris = scene.n_triangles

    band_rows = 1
    peer = None
    if world > 1:
        band_rows = band_rows_for(scene.height, world)
        api.swglSetStripe(rank, world, band_rows)
        peer = multigpu.PeerColorTarget(api, dist, rank, world)

    stream = torch.cuda.ExternalStream(api.swglGetStream(), device=torch.device("cuda", local))
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.int32, device=f"cuda:{local}")

    def frame():
        api.glClear(G.GL_COLOR_BUFFER_BIT | G.GL_DEPTH_BUFFER_BIT)
        api.glDrawElements(G.GL_TRIANGLES, n_draw, G.GL_UNSIGNED_INT, None)

    def barrier():
        api.swglFinish()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()
    launches0 = api.swglGetOption(b"kernel_launches")

    # ---- value: device-resident frames, per-step CUDA events, L2 flushed between steps ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        with torch.cuda.stream(stream):
            flush.zero_()
        a.record(stream)
        frame()
        b.record(stream)
    barrier()
    launches = api.swglGetOption(b"kernel_launches") - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    if dist is not None:
        t = torch.tensor([total_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n_tris / (ms_per_step * 1e-3)

    stats = sw.swglStats()
    api.swglGetStats(C.byref(stats))
    sd = stats.as_dict()
    shaded, tested = sd["shaded"], sd["tested"]
    if dist is not None:
        t = torch.tensor([shaded, tested], device=f"cuda:{local}", dtype=torch.int64)
        dist.all_reduce(t)
        shaded, tested = int(t[0].item()), int(t[1].item())

    # ---- roofline of the dominant kernel: per-kernel CUDA events inside the library ----
    api.swglSetOption(b"stage_timing", 1)
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        frame()
    api.swglFinish()
    nd = max(1, api.swglGetOption(b"stage_draws"))
    stage_us = {n: api.swglGetOption(f"stage_ns_{i}".encode()) / 1e3 / nd for i, n in enumerate(STAGES)}
    api.swglSetOption(b"stage_timing", 0)
    peak, peak_src = measured_peaks()
    dom = max(stage_us, key=stage_us.get)
    fb_bytes = scene.width * scene.height * 8 // world
    dom_bytes = fb_bytes if dom == "k_raster" else scene.algorithmic_bytes()
    achieved = dom_bytes / (stage_us[dom] * 1e-6) / 1e9
    # per-launch counters of the dominant kernel from the committed ncu capture of this config (stamped with the
    # commit it was taken at): DRAM traffic, and the warp-instruction count behind the SM-issue roofline
    ctr, captured_at = ncu_counters(args.config, dom) if world == 1 else (None, None)
    traffic = ctr.get("dram_bytes") if ctr else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes,
                "kernel_us": stage_us[dom], "stage_us": stage_us,
                "frame_algorithmic_bytes": scene.algorithmic_bytes(),
                "frame_frac": scene.algorithmic_bytes() / (ms_per_step * 1e-3) / 1e9 / peak,
                "frame_frac_nominal_8TBs": scene.algorithmic_bytes() / (ms_per_step * 1e-3) / 8.0e12}
    if ctr and ctr.get("warp_inst"):
        # what actually bounds the kernel: warp instructions per launch (ncu) / its live duration against the
        # issue rate of 148 SMs x 4 schedulers x 1 instruction per clock at the SM clock sampled below
        roofline["issue"] = {"warp_inst_per_launch": int(ctr["warp_inst"]), "kernel_us": stage_us[dom],
                             "achieved_ginst_s": ctr["warp_inst"] / (stage_us[dom] * 1e-6) / 1e9,
                             "ncu_issue_active_pct": ctr.get("issue_active_pct"), "ncu_dram_pct": ctr.get("dram_pct"),
                             "source": f"profiles/ncu_counters.json, captured at {captured_at}"}

    # ---- N > 1: the image assembled on rank 0 must equal the unsharded render, bit for bit ----
    def assembled_equals_single():
        barrier()
        frame()
        barrier()
        out = None
        if rank == 0:
            multi = G.frame_color(api, scene.width, scene.height)
            api.swglSetStripe(0, 1, 1)
            frame()
            single = G.frame_color(api, scene.width, scene.height)
            api.swglSetStripe(rank, world, band_rows)
            out = {"equal_to_single_gpu": bool(np.array_equal(multi, single)),
                   "mismatching_pixels": int((multi != single).sum())}
        barrier()
        return out

    mg_check = assembled_equals_single() if world > 1 else None

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ----
    # N > 1 (default --e2e-mode group): what a C application does -- ONE host thread, swglSetDeviceCount(N) before
    # glInit, then the reference's own calls.  Rank 0 drives all N GPUs of the box through the library's device
    # group (uploads split over the N PCIe links and completed over NVLink, frame assembled in one pinned mirror);
    # the other ranks wait on the rendezvous store without touching their GPUs.
    e2e = None
    if world > 1 and args.e2e_mode == "group":
        peer.close()
        peer = None
        barrier()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                e2e = group_e2e(api, sw, G, scene, args, world, local)
            except Exception as exc:          # report, do not hang the other ranks
                e2e = {"error": repr(exc)}
            store.set("swgl_b200_e2e_done", "1")
        else:
            store.wait(["swgl_b200_e2e_done"])
            e2e = {}                          # rank 0's record is the one that is printed
        barrier()
        if rank == 0:
            api.swglSetDeviceCount(1)         # nothing below renders at N > 1
            api.swglSetDevice(local)
    if e2e is None:
        e2e = ranks_e2e(api, G, multigpu, dist, torch, scene, args, rank, world, local, frame, barrier, assembled_equals_single, peer)
        peer = None
    verts = idx = None
    if world == 1:
        verts = multigpu.HostArray(api, scene.vertices)
        idx = multigpu.HostArray(api, scene.indices)

    # ---- the same step pipelined: two geometry sets and two frame mirrors; step N's
    # upload overlaps frame N-1 on the device, its frame is collected during step N+1 ----
    e2e_pipe = None
    if world == 1:
        sets = [G.add_geometry(api, scene, indexed=True, named=True)[:3] for _ in range(2)]

        def pipe_step(i, prev):
            vao, vbo, ebo = sets[i & 1]
            api.glBindVertexArray(vao)
            api.glBindBuffer(G.GL_ARRAY_BUFFER, vbo)
            api.glBindBuffer(G.GL_ELEMENT_ARRAY_BUFFER, ebo)
            api.swglBufferRespecify(G.GL_ARRAY_BUFFER, verts.nbytes, C.c_void_p(verts.ptr))
            api.swglBufferRespecify(G.GL_ELEMENT_ARRAY_BUFFER, idx.nbytes, C.c_void_p(idx.ptr))
            frame()
            t = api.swglFrameSubmit()
            if prev:
                api.swglFrameWait(prev)
            return t

        prev = 0
        for i in range(3):
            prev = pipe_step(i, prev)
        api.swglFrameWait(prev)
        barrier()
        t0 = time.perf_counter()
        prev = 0
        for i in range(args.steps):
            prev = pipe_step(i, prev)
        api.swglFrameWait(prev)
        pipe_s = (time.perf_counter() - t0) / args.steps
        barrier()
        e2e_pipe = {"value": n_tris / pipe_s, "unit": METRIC, "ms_per_step": pipe_s * 1e3,
                    "note": "swglFrameSubmit/swglFrameWait, two geometry sets: same bytes per step as e2e, each frame is read one step later"}

    clocks = sampler.stop() if rank == 0 else {}   # sampled across the value, roofline and e2e legs
    if "issue" in roofline:
        mhz = clocks.get("sm_mhz") or 1965.0
        roofline["issue"]["peak_ginst_s"] = 148 * 4 * mhz * 1e6 / 1e9
        roofline["issue"]["frac"] = roofline["issue"]["achieved_ginst_s"] / roofline["issue"]["peak_ginst_s"]

    # ---- parity of the timed workload against the reference's known-answer frame (N = 1; at N > 1 the assembled
    # image has been compared with the single-GPU one above) ----
    parity = None
    if world == 1:
        frame()
        st1 = sw.swglStats()
        api.swglGetStats(C.byref(st1))
        parity = parity_block(api, scene, args.config, st1.as_dict())

    # ---- CPU baseline: the unmodified reference on one host core, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_sample(scene, target_seconds=12.0)

    if verts is not None:
        verts.free()
        idx.free()
    if peer is not None:
        peer.close()
    err = api.swglGetLastError().decode()

    # ---- the other BASELINE configs (N = 1): frame time, stages, roofline fraction, parity ----
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        for cfg in (1, 2, 3, 5):
            if cfg == args.config:
                continue
            configs[f"C{cfg}"] = measure_config(
                api, sw, G, S, torch, cfg, max(5, args.steps // 2),
                lambda: torch.cuda.ExternalStream(api.swglGetStream(), device=torch.device("cuda", local)), flush, peak)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.config, scene, world),
            "shaded_fragments_per_s": shaded / (ms_per_step * 1e-3),
            "tested_fragments_per_s": tested / (ms_per_step * 1e-3),
            "frame_ms": ms_per_step, "shaded_fragments": shaded, "tested_fragments": tested,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
        }
        if e2e_pipe is not None:
            line["e2e_pipelined"] = e2e_pipe
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if mg_check is not None:
            line["multi_gpu_check"] = mg_check
        if parity is not None:
            line["parity"] = parity
        if configs is not None:
            line["configs"] = configs
        if err:
            line["error"] = err
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _timed_cpu(scene, tris, ref, O):
    import copy

    sample = copy.copy(scene)
    if scene.indices is not None:
        sample.indices = scene.indices[: 3 * tris]
    else:
        sample.vertices = scene.vertices[: 3 * tris]
    if ref is not None:
        return ref.timed_frame(sample)
    dt, _ = O.Restatement().timed_frame(sample)
    return dt, None


class pinned_to_one_core:
    """The reference is single-threaded: run it on one host core, restore the mask afterwards."""

    def __enter__(self):
        self.mask = None
        try:
            self.mask = os.sched_getaffinity(0)
            core = max(self.mask)            # away from core 0, where interrupt handling tends to land
            os.sched_setaffinity(0, {core})
            return core
        except (AttributeError, OSError):
            return None

    def __exit__(self, *exc):
        if self.mask:
            try:
                os.sched_setaffinity(0, self.mask)
            except OSError:
                pass
        return False


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_sample(scene, target_seconds: float, tris: int | None = None, reps: int | None = None):
    """Time the compiled reference (oracle/_ref; else the C port) on one host core.

    The sample is the first `tris` triangles of the workload drawn into the full-size
    framebuffer (glClear + glDrawArrays), repeated `reps` times; both are sized from a short
    calibration draw so the whole call costs about `target_seconds` of CPU work."""
    from oracle import pyoracle as O

    kind = "reference"
    try:
        ref = O.Reference()
    except Exception:
        ref = None
        kind = "port"
    n_tris = scene.n_triangles
    if tris is None:
        cal = min(n_tris, 20000)
        dt, _ = _timed_cpu(scene, cal, ref, O)
        rate = cal / max(dt, 1e-6)
        tris = int(min(n_tris, max(2000, target_seconds * rate)))
        reps = max(1, int(round(target_seconds / (tris / rate)))) if reps is None else reps
    reps = reps or 1
    times, clear_s = [], None
    with pinned_to_one_core() as core:
        for _ in range(reps):
            dt, clear_s = _timed_cpu(scene, tris, ref, O)
            times.append(dt)
    dt = statistics.mean(times)
    return {"value": tris / dt, "unit": METRIC, "cores": 1, "kind": kind, "core": core, "cpu_model": cpu_model(),
            "sample": f"first {tris} of {n_tris} triangles of the same scene, full-size framebuffer, "
                      f"glClear+glDrawArrays x{reps} (mean {dt:.2f} s per frame, clear {clear_s if clear_s is None else round(clear_s, 3)} s)",
            "seconds": dt, "tris": tris, "reps": reps, "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from swgl_b200 import scenes as S

    scene = S.config(args.config)
    total = max(1, args.steps + args.warmup)
    cal = cpu_reference_sample(scene, 0.0, tris=min(scene.n_triangles, 20000), reps=1)
    budget = 90.0 / total                                   # whole run about 1.5 minutes
    tris = int(min(scene.n_triangles, max(2000, budget * cal["value"])))
    res = None
    times = []
    for i in range(total):
        res = cpu_reference_sample(scene, 0.0, tris=tris, reps=1)
        if i >= args.warmup:
            times.append(res["seconds"])
    dt = statistics.mean(times)
    value = tris / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, scene, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": 1, "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--config", type=int, default=4, help="BASELINE.json config 1..5 (default 4: 4K, 1M triangles)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-mode", default="group", choices=["group", "ranks"],
                    help="N > 1 end-to-end step: rank 0 drives all GPUs through swglSetDeviceCount (group), or one process per GPU (ranks)")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config block (C1, C2, C3, C5) of the N=1 line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
