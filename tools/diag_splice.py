#!/usr/bin/env python
"""Where does the time of an assembly ON TOP of a resident matrix go?  (development tool)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
import xsparse_b200 as xsb

mesh = 128
n = mesh ** 3
h = xsb.Handle(n, n)
h.emit_p1fem(mesh, mesh, mesh, flavour=xsb.RAW)
h.flush()
for it in range(4):
    t0 = time.perf_counter()
    h.timer_start()
    h.emit_p1fem(mesh, mesh, mesh, flavour=xsb.RAW)
    ms_emit = h.timer_stop()
    t1 = time.perf_counter()
    h.timer_start()
    h.flush()
    ms_flush = h.timer_stop()
    t2 = time.perf_counter()
    print(f"it {it}: emit {ms_emit:.3f} ms (host {1e3*(t1-t0):.3f}), flush {ms_flush:.3f} ms (host {1e3*(t2-t1):.3f}) path {h.flush_stats()['column_path']} launches {h.flush_stats()['kernel_launches']}")
