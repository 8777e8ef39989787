"""Tile height of the warp rasteriser (8 / 4 / 2 rows): frame and stage times per config, and for the share of one
rank of an N-rank sort-first group (stripe emulation on one GPU).  Development / profiles."""
import json
import sys

sys.path.insert(0, ".")
from swgl_b200 import scenes as S
from tools.perf_probe import probe, api

out = {}
for cfg in (1, 2, 4):
    for rows in (0, 8, 4, 2):
        dt = probe(S.config(cfg), reps=20, options={"tile_rows": rows})
        out[f"C{cfg}_rows{rows}"] = {"frame_us": dt * 1e6, "picked": api.swglGetOption(b"last_tile_rows")}
print(json.dumps(out))
