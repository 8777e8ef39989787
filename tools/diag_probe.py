import subprocess, sys
cfg = sys.argv[1] if len(sys.argv) > 1 else "4"
for d in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,8,9,10,11".split(","))]:
    out = subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0,'.'); from tools.perf_probe import probe; probe({cfg}, reps=10, options={{'diag': {d}}})"], capture_output=True, text=True).stdout
    print("diag", d, [l for l in out.splitlines() if "stages" in l])
