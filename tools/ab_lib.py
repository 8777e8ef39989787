"""A/B of two builds of libswgl_b200.so on one GPU box (development).

    python tools/ab_lib.py <cfg> tools/ab/prev.so [more.so ...]   # "cur" = the in-tree build
"""
import os, subprocess, sys
cfg = sys.argv[1]
libs = [("cur", "")] + [(os.path.basename(p), os.path.abspath(p)) for p in sys.argv[2:]]
for rnd in range(3):
    for name, path in libs:
        env = dict(os.environ)
        if path:
            env["SWGL_B200_LIB"] = path
        out = subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0,'.'); from tools.perf_probe import probe; probe({cfg}, reps=20)"],
                             capture_output=True, text=True, env=env)
        print(name, [l.strip() for l in out.stdout.splitlines() if "stages" in l], out.stderr[-200:] if out.returncode else "")
