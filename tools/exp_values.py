#!/usr/bin/env python
"""Values-only re-assembly of fdrand n^3 into the frozen pattern (development tool)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
import xsparse_b200 as xsb

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = nx ** 3
g = xsb.Handle(n, n)
g.set_precount(False)
g.emit_fdrand(nx, nx, nx, seed=100)
cnt = g.pending
dI = torch.empty(cnt, dtype=torch.int64, device="cuda"); dJ = torch.empty_like(dI)
dV = torch.empty(cnt, dtype=torch.float64, device="cuda")
got = C.c_int64(0)
c = xsb.capi
c.check(c.lib().xsb_debug_fetch_staged(g._h, 0, dI.data_ptr(), dJ.data_ptr(), dV.data_ptr(), None, cnt, C.byref(got)), g._h)
g.set_precount(True)
nnz, _ = g.flush()
g.freeze_pattern(dI, dJ, count=cnt)
for name, m in (("deterministic", xsb.DETERMINISTIC), ("fast", xsb.FAST)):
    for _ in range(2):
        g.reassemble_values(dV, m, count=cnt, zero_first=True)
    g.synchronize(); g.timer_start()
    for _ in range(10):
        g.reassemble_values(dV, m, count=cnt, zero_first=True)
    ms = g.timer_stop() / 10
    b = 12 * cnt + 8 * nnz
    print(f"{name}: {ms:.4f} ms  {b/ms/1e6:.0f} GB/s algorithmic")
