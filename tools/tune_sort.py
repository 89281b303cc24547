"""Micro-benchmark of the onesweep variants (run on the GPU box): per-pass time on random keys."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

ge.build()
import xsparse_b200 as xsb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nbits = int(sys.argv[2]) if len(sys.argv) > 2 else 42
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(7))
h = xsb.Handle(16, 16)
for v in variants:
    r = h.sort_selftest(n, nbits, v, reps=3)
    gbs = 32 * n / (r["ms_per_pass"] * 1e-3) / 1e9
    print(json.dumps({"variant": v, "n": n, "nbits": nbits, **r, "GBps_per_pass": round(gbs, 1)}))
