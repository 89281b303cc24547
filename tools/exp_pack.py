#!/usr/bin/env python
"""Throughput of the generic insertion path with DEVICE-resident input (development tool):
xsb_insert_triplets / xsb_insert_batch of the FEM 128^3 stream, then flush."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
import xsparse_b200 as xsb

mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = mesh ** 3
g = xsb.Handle(n, n)
g.set_precount(False)
g.emit_p1fem(mesh, mesh, mesh, flavour=xsb.RAW)
cnt = g.pending
dI = torch.empty(cnt, dtype=torch.int64, device="cuda"); dJ = torch.empty_like(dI)
dV = torch.empty(cnt, dtype=torch.float64, device="cuda")
got = C.c_int64(0)
c = xsb.capi
c.check(c.lib().xsb_debug_fetch_staged(g._h, 0, dI.data_ptr(), dJ.data_ptr(), dV.data_ptr(), None, cnt, C.byref(got)), g._h)
g.close()
dT = torch.empty((cnt, 2), dtype=torch.int64, device="cuda")
dT[:, 0] = dI | (dJ << 32)
dT[:, 1] = dV.view(torch.int64)
torch.cuda.synchronize()
h = xsb.Handle(n, n)
h.set_profiling(True)
for rep in range(4):
    h.reset()
    h.timer_start()
    h.insert_triplets(dT, xsb.RAW, 0, cnt)
    ms_t = h.timer_stop()
    h.flush()
    st = h.flush_stats()
    print(f"triplets: insert {ms_t:.3f} ms  flush {st['ms_total']:.3f} ms path {st['column_path']} nnz {st['nnz_new']}")
for rep in range(3):
    h.reset()
    h.timer_start()
    h.insert_batch(dI, dJ, dV, xsb.RAW)
    ms_t = h.timer_stop()
    h.flush()
    st = h.flush_stats()
    print(f"(I,J,V): insert {ms_t:.3f} ms  flush {st['ms_total']:.3f} ms path {st['column_path']}")
