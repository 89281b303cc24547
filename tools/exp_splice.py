#!/usr/bin/env python
"""Stage times of repeated 1 % splices into the resident FEM matrix (development tool)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
ge.build()
import xsparse_b200 as xsb

mesh = 128
n = mesh ** 3
h = xsb.Handle(n, n)
h.emit_p1fem(mesh, mesh, mesh, flavour=xsb.RAW)
h.flush()
rng = np.random.default_rng(5)
k = h.nnz // 100
h.set_profiling(True)
for it in range(5):
    I = rng.integers(1, n + 1, k); J = rng.integers(1, n + 1, k); V = rng.standard_normal(k)
    h.insert_batch(I, J, V, xsb.UPDATE)
    h.timer_start()
    r = h.flush()
    ms = h.timer_stop()
    st = h.flush_stats()
    print(it, f"wall {ms:.3f}", r, {k2: (round(v, 3) if isinstance(v, float) else v) for k2, v in st.items() if v})
