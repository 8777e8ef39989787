"""Summarise an ncu --set full report: one row per kernel launch (markdown) + traffic json.

usage: python tools/ncu_summary.py <report.ncu-rep> <out.md> [ncu_counters.json config-key commit]

The counters file feeds bench.py's `roofline.traffic` and `roofline.issue` (per-launch DRAM bytes and warp
instructions of each kernel), stamped with the commit the capture was taken at.
"""
import csv, json, subprocess, sys, os

rep, out_md = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time us"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
]
lines = ["| " + " | ".join(c[1] for c in cols) + " |", "|" + "---|" * len(cols)]
traffic = {}
for r in rows[2:]:
    vals = []
    for key, _ in cols:
        v = r[ix[key]] if key in ix else ""
        try:
            f = float(v)
            v = f"{f:.2f}" if abs(f) < 1e6 and f != int(f) else f"{int(f)}"
        except ValueError:
            v = v.replace("(DrawParams)", "").replace("void ", "")
        vals.append(v)
    lines.append("| " + " | ".join(vals) + " |")
    name = r[ix["Kernel Name"]].replace("void ", "").split("<")[0].split("(")[0]
    rd = float(r[ix["dram__bytes_read.sum"]]); wr = float(r[ix["dram__bytes_write.sum"]])
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    rdb = rd * scale.get(units[ix["dram__bytes_read.sum"]], 1.0)
    wrb = wr * scale.get(units[ix["dram__bytes_write.sum"]], 1.0)
    def num(key):
        try:
            return float(r[ix[key]])
        except (KeyError, ValueError):
            return None
    traffic["k_raster" if name.startswith("k_raster") else name] = {
        "dram_bytes": int(rdb + wrb), "warp_inst": int(num("smsp__inst_executed.sum") or 0),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "threads_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "ncu_time_us": num("gpu__time_duration.sum")}
with open(out_md, "w") as f:
    f.write(f"# ncu --set full summary of `{os.path.basename(rep)}`\n\n")
    f.write("One frame of the workload (glClear fused + glDrawElements); per-launch values, cold-cache and\n"
            "serialised by the profiler: compare SHARES, not absolutes.\n\n")
    f.write("\n".join(lines) + "\n")
if len(sys.argv) > 4:
    path, key = sys.argv[3], sys.argv[4]
    cur = {}
    if os.path.exists(path):
        cur = json.load(open(path))
    cur.setdefault(key, {}).update(traffic)
    if len(sys.argv) > 5:
        cur["captured_at"] = sys.argv[5]
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
print("\n".join(lines))
