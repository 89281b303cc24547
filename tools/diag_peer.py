#!/usr/bin/env python
"""Repro: idle ranks in a sparse peer exchange (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
ge.build()
import xsparse_b200 as xsb
world, m, n = 4, 200, 300
splits = [0, 75, 150, 225, 300]
handles = [xsb.Handle(m, n, slab=(world, r, splits)) for r in range(world)]
caps = [[0 if s == d or abs(s - d) > 1 else 4000 for s in range(world)] for d in range(world)]
for h in handles: h.peer_exchange_create(caps)
for h in handles: h.peer_exchange_connect_local(handles)
rng = np.random.default_rng(1)
handles[0].insert_batch(rng.integers(1, m + 1, 500), rng.integers(76, 151, 500), rng.standard_normal(500), xsb.RAW)
for h in handles: h.route_pack_peer()
for h in handles: h.route_unpack_peer()
for r, h in enumerate(handles):
    try:
        print(r, h.flush())
    except Exception as e:
        print(r, "ERR", e)

def step(tag):
    for h in handles: h.route_pack_peer()
    for h in handles: h.route_unpack_peer()
    for r, h in enumerate(handles):
        try:
            print(tag, r, h.flush(), h.flush_stats()["column_path"])
        except Exception as e:
            print(tag, r, "ERR", str(e)[:60])
handles[0].insert_batch(rng.integers(1, m + 1, 5000), rng.integers(76, 151, 5000), rng.standard_normal(5000), xsb.RAW)
step("overflow")
for h in handles: h.reset()
handles[0].insert_batch(np.array([1]), np.array([n]), np.array([1.0]), xsb.RAW)
step("noblock")
for h in handles: h.reset()
step("empty")
