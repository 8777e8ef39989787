"""Aggregate tools/ncu_lines.py output for k_raster_warp into phases (line ranges found from markers in the source)."""
import re, subprocess, sys, os
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "_Z13k_raster_warpILi1EEv10DrawParams"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "swgl_b200/csrc/swgl_raster_warp.cuh")).read().splitlines()
KSTART = next(i for i, l in enumerate(src) if "k_raster_warp(const __grid_constant__" in l)
def find(marker):
    return next(i + 1 for i, l in enumerate(src) if i >= KSTART and marker in l)
marks = [("helpers(sort32,merge,stage)", 1), ("prologue+stage", find("k_raster_warp(const __grid_constant__")), ("sort", find("ascending primitive id = submission order ----")),
         ("phaseA", find("for (uint32_t base = 0; base < n_list; base += take)")), ("scan+spans", find("---- S: one exclusive scan")),
         ("locate", find("---- phase B: lane = fragment")), ("fetch+weights", find("FragIn fi;")),
         ("shade", find("the fragment shader does not read the framebuffer")), ("commit", find("---- ordered commit")),
         ("writeback", find("---- write-back"))]
math = open(os.path.join(root, "swgl_b200/csrc/swgl_dev_math.cuh")).read().splitlines()
def mfind(marker):
    return next(i + 1 for i, l in enumerate(math) if marker in l)
mmarks = [("math misc", 1), ("fdiv", mfind("float fdiv(float x, float y)") - 5), ("div_shared", mfind("Division with a shared, refined reciprocal")), ("phaseA(math)", mfind("float canon_nan") - 3), ("bary_setup", mfind("struct BaryConst") - 1),
          ("weights slow", mfind("Barycentric + perspective correction + depth")), ("weights fast", mfind("the same eight divisions with shared")), ("prim_consts", mfind("Per-primitive constants of the fragment arithmetic")), ("weights fast", mfind("frag_weights() from staged constants")), ("blend etc", mfind("clamp, unpack destination, blend, pack"))]
env = dict(os.environ, NCU_KERNEL="k_raster_warp")
out = subprocess.run([sys.executable, os.path.join(root, "tools/ncu_lines.py"), rep, kern, "3000"], capture_output=True, text=True, env=env).stdout
cats = {}
for line in out.splitlines():
    m = re.match(r'\s*([\d.]+)% inst\s+([\d.]+)% smp\s+thr/inst\s+([\d.]+)\s+(\S+):(\d+)', line)
    if not m: continue
    p, s, f, l = float(m.group(1)), float(m.group(2)), m.group(4), int(m.group(5))
    if f == "swgl_raster_warp.cuh":
        c = [n for n, a in marks if a <= l][-1]
    elif f == "swgl_dev_math.cuh":
        c = [n for n, a in mmarks if a <= l][-1]
    elif f == "swgl_raster_frag.cuh":
        c = "clamp" if l <= 94 else "blend"
    elif f == "swgl_dev.cu":
        c = "dev.cu (shade/walk_to_row)"
    else:
        c = f
    a = cats.setdefault(c, [0, 0]); a[0] += p; a[1] += s
print(out.splitlines()[0] if out else "")
for k, v in sorted(cats.items(), key=lambda kv: -kv[1][0]):
    print('%6.1f%% inst %6.1f%% smp  %s' % (v[0], v[1], k))
