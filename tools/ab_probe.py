"""A/B of runtime options on one box: python tools/ab_probe.py <cfg> opt=v0,v1 [reps]"""
import sys
sys.path.insert(0, ".")
from tools.perf_probe import probe
cfg = int(sys.argv[1])
name, vals = sys.argv[2].split("=")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
for rnd in range(3):
    for v in vals.split(","):
        print(f"--- {name}={v}")
        probe(cfg, reps=reps, options={name: int(v)})
