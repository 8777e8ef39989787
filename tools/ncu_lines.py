"""Join an ncu SASS source page (csv) with nvdisasm -g line info: per-source-line instruction counts.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel mangled name> [top N]
"""
import csv, re, subprocess, sys, os, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "swgl_b200", "libswgl_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if "sm_100a" in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout.splitlines()
# slice kernel
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + kern + ":"))
end = next((i for i in range(start + 1, len(sass)) if sass[i].startswith("\t.section")), len(sass))
cur = ("?", 0)
ins_lines = []   # per instruction: (file, line)
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l) or re.match(r"\s+[A-Z@!]", l):
        if re.match(r"\s+\.", l):
            continue
        ins_lines.append((cur, l.strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
print(f"sass instructions: nvdisasm {len(ins_lines)} ncu {len(data)}", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
n = min(len(ins_lines), len(data))
for k in range(n):
    (fl, ln), _ = ins_lines[k]
    r = data[k]
    agg[(fl, ln)][0] += int(r[ix["Instructions Executed"]] or 0)
    agg[(fl, ln)][1] += int(r[ix["Thread Instructions Executed"]] or 0)
    agg[(fl, ln)][2] += int(r[ix["# Samples"]] or 0)
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
srcs = {}
print(f"total warp-instr {tot}  samples {tots}")
for (fl, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    path = os.path.join(os.path.dirname(lib), "csrc", fl)
    if fl not in srcs:
        try:
            srcs[fl] = open(path).read().splitlines()
        except OSError:
            srcs[fl] = []
    text = srcs[fl][ln - 1].strip()[:90] if 0 < ln <= len(srcs[fl]) else ""
    print(f"{100*v[0]/tot:5.1f}% inst {100*v[2]/tots:5.1f}% smp  thr/inst {v[1]/max(v[0],1):5.1f}  {fl}:{ln}  {text}")

# ---- region summary for swgl_raster_frag.cuh (line ranges of the kernel's phases) ----
regions = [("kernel prologue/list", "swgl_raster_frag.cuh", 106, 123), ("stage tile", "swgl_raster_frag.cuh", 124, 168), ("sort", "swgl_raster_frag.cuh", 169, 216),
           ("phase A1", "swgl_raster_frag.cuh", 217, 256), ("phase A2 walk", "swgl_raster_frag.cuh", 257, 280), ("A scan+map", "swgl_raster_frag.cuh", 281, 321),
           ("B locate", "swgl_raster_frag.cuh", 322, 352), ("B early shade", "swgl_raster_frag.cuh", 353, 371), ("B commit", "swgl_raster_frag.cuh", 372, 412),
           ("write-back", "swgl_raster_frag.cuh", 413, 446), ("stats", "swgl_raster_frag.cuh", 447, 470), ("blend", "swgl_raster_frag.cuh", 85, 105),
           ("scan util", "swgl_raster_frag.cuh", 64, 84)]
if "warp" in kern:
    regions = [("prologue+list", "swgl_raster_warp.cuh", 120, 153), ("stage tile", "swgl_raster_warp.cuh", 154, 169), ("stage tile", "swgl_raster_warp.cuh", 83, 119),
               ("sort", "swgl_raster_warp.cuh", 170, 185), ("sort", "swgl_raster_warp.cuh", 42, 82), ("phase A", "swgl_raster_warp.cuh", 186, 222),
               ("scan", "swgl_raster_warp.cuh", 223, 231), ("B locate", "swgl_raster_warp.cuh", 232, 268), ("B weights+shade", "swgl_raster_warp.cuh", 269, 288),
               ("B commit", "swgl_raster_warp.cuh", 289, 327), ("write-back", "swgl_raster_warp.cuh", 328, 362), ("stats", "swgl_raster_warp.cuh", 363, 380),
               ("blend", "swgl_raster_frag.cuh", 85, 115)]
summ = collections.defaultdict(lambda: [0, 0])
for (fl, ln), v in agg.items():
    name = None
    for r in regions:
        if fl == r[1] and r[2] <= ln <= r[3]:
            name = r[0]
    if name is None:
        name = fl
    summ[name][0] += v[0]; summ[name][1] += v[2]
print("---- regions ----")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*v[0]/tot:5.1f}% inst {100*v[1]/tots:5.1f}% smp  {k}")
