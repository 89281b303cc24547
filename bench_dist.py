"""N>1 arm of bench.py: weak scaling of the partitioned FEM assembly over column-owning ranks.

Global mesh: 128 x 128 x (127*N + 1) nodes (Kuhn 6-tet).  Rank r emits the tetrahedra of its
127 cube layers (245 805 960 rawupdateindex! calls, the same per-GPU work as the N=1 workload)
and owns the columns of the z-planes [127 r, 127 (r+1)) (the last rank also owns the top plane).
Records whose column lies on the interface plane travel to the rank above in one NCCL
all-to-all-v; every rank then merges into its own CSC slab.  The assembled matrix is left
sharded (slab-local colptr + global offset); that is what is timed.
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist


def run(args, xsb, rank, world, local):
    import bench

    # watchdog: a rank stuck in a collective dumps its Python stack and exits instead of hanging the box
    import faulthandler
    import sys

    faulthandler.dump_traceback_later(int(os.environ.get("XSB_BENCH_WATCHDOG_S", "240")), exit=True, file=sys.stderr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from xsparse_b200 import dist as xd

    dev = torch.device("cuda", local)
    if getattr(args, "workload", "fem") == "fd400":
        return run_fd(args, xsb, xd, bench, rank, world, local, dev)
    nx = ny = args.mesh
    layers = args.mesh - 1
    nz_nodes = layers * world + 1
    N = nx * ny * nz_nodes
    plane = nx * ny
    splits = [plane * layers * r for r in range(world)] + [N]
    mode = xsb.DETERMINISTIC if args.mode == "deterministic" else xsb.FAST
    D = xd.DistExtendableSparseMatrix(N, N, splits=splits, device=local)
    h = D.h
    h.set_profiling(True)
    n_ins_rank = xsb.capi.stream_count_p1fem(nx, ny, layers + 1)

    def step():
        h.reset()
        h.emit_p1fem(nx, ny, nz_nodes, flavour=xsb.RAW, cz_range=(layers * rank, layers * (rank + 1)))
        return D.flush(mode, wait=False)  # the 16-byte offsets all-gather is launched, not waited for

    for _ in range(args.warmup):
        step()
    h.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = h.kernel_launches
    stage = {}
    with bench.ClockSampler(local) as clk:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        h.timer_start()  # CUDA events on the stream the library launches on
        for _ in range(args.steps):
            nnz, _ = step()
            st = h.flush_stats()
            for k, v in st.items():
                if k.startswith("ms_"):
                    stage[k] = stage.get(k, 0.0) + v
        ms_local = h.timer_stop()
        torch.cuda.synchronize()
        dist.barrier()
        launches_timed = h.kernel_launches - launches0
        # max over ranks, on the device clock
        tmax = torch.tensor([ms_local], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
        # keep the clock sampler busy for about half a second: the SAME number of extra steps on every rank
        # (a step holds collectives; a per-rank time limit would let the ranks disagree and hang)
        n_fill = int(max(0.0, 500.0 - ms) / max(ms / args.steps, 1e-3)) + 1
        for _ in range(min(n_fill, 200)):
            step()
        torch.cuda.synchronize()
        dist.barrier()
    launches = torch.tensor([launches_timed], dtype=torch.int64, device=dev)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    st = h.flush_stats()
    ms_step = ms / args.steps
    value = world * n_ins_rank / (ms_step / 1e3)
    peak, peak_src = bench.peaks()
    roof = bench.roofline(st, {k: v / args.steps for k, v in stage.items()}, None, h.n, peak, peak_src, None)
    roof["kernel"] += " [rank 0]"

    e2e = measure_e2e(args, xsb, xd, rank, world, local, mode)
    clocks = clk.summary()
    if rank == 0:
        cfg = bench.workload(args)
        cfg["workload"] = (f"P1-FEM Laplacian+mass, {nx}x{ny}x{nz_nodes}-node Kuhn mesh sharded over {world} ranks "
                           f"({layers} cube layers = {n_ins_rank} rawupdateindex! calls per rank), column-slab ownership, "
                           f"NCCL all-to-all-v of the interface plane, CSC left sharded")
        cfg["parallelism"] = f"column-slab x{world}"
        cfg["exchange"] = dict(D.last_exchange)
        cfg["host_phase_ms_last_step"] = dict(zip(["route_count", "copy_out", "all_to_all", "append", "flush", "offsets"],
                                                  [round(x, 3) for x in D.last_phase_ms]))
        line = {
            "metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": roof,
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches.item()), "clocks": clocks,
            "nnz_global": int(D.nnz_global), "n_inserted": int(world * n_ins_rank),
        }
        print(json.dumps(line))
    h.close()
    dist.barrier()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


def measure_e2e(args, xsb, xd, rank, world, local, mode):
    """End to end at N ranks with HOST buffers: every step each rank copies its insertion stream (16-byte
    triplets) from pinned host memory, inserts, routes, flushes and reads its CSC slab back to pinned host memory."""
    import ctypes as C

    emesh = args.e2e_mesh
    layers = emesh - 1
    nz_nodes = layers * world + 1
    N = emesh * emesh * nz_nodes
    splits = [emesh * emesh * layers * r for r in range(world)] + [N]
    D = xd.DistExtendableSparseMatrix(N, N, splits=splits, device=local)
    g = D.h
    g.set_precount(False)  # the host array holds the stream in call order
    g.emit_p1fem(emesh, emesh, nz_nodes, flavour=xsb.RAW, cz_range=(layers * rank, layers * (rank + 1)))
    cnt = g.pending
    dI = torch.empty(cnt, dtype=torch.int64, device="cuda")
    dJ = torch.empty(cnt, dtype=torch.int64, device="cuda")
    dV = torch.empty(cnt, dtype=torch.float64, device="cuda")
    got = C.c_int64(0)
    c = xsb.capi
    c.check(c.lib().xsb_debug_fetch_staged(g._h, 0, dI.data_ptr(), dJ.data_ptr(), dV.data_ptr(), None, cnt,
                                           C.byref(got)), g._h)
    # the stream as 16-byte triplets {u32 row, u32 col, f64 val} (xsb_insert_triplets), as at N=1
    dT = torch.empty((cnt, 2), dtype=torch.int64, device="cuda")
    dT[:, 0] = dI | (dJ << 32)
    dT[:, 1] = dV.view(torch.int64)
    hT = torch.empty((cnt, 2), dtype=torch.int64, pin_memory=True).copy_(dT)
    torch.cuda.synchronize()
    del dI, dJ, dV, dT
    g.reset()
    g.set_precount(True)
    g.insert_triplets(hT, xsb.RAW, 0, cnt)
    nnz, _ = D.flush(mode)
    ocp = torch.empty(g.n + 1, dtype=torch.int64, pin_memory=True)
    orv = torch.empty(nnz, dtype=torch.int64, pin_memory=True)
    onz = torch.empty(nnz, dtype=torch.float64, pin_memory=True)

    def step():
        g.reset()
        g.insert_triplets(hT, xsb.RAW, 0, cnt)
        D.flush(mode, wait=False)
        g.fetch_csc(ocp, orv, onz)

    step()
    steps = max(1, min(args.steps, 3))
    torch.cuda.synchronize()
    dist.barrier()
    g.timer_start()
    for _ in range(steps):
        step()
    ms_e2e = g.timer_stop()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ms_e2e / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([cnt, 16 * cnt, 8 * (g.n + 1) + 16 * int(nnz)], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    g.close()
    return {"value": int(tot[0].item()) / (ms / 1e3), "unit": "entries/s", "h2d_bytes_per_step": int(tot[1].item()),
            "d2h_bytes_per_step": int(tot[2].item()), "ms_per_step": ms,
            "workload": f"P1-FEM {emesh}x{emesh}x{nz_nodes}-node mesh over {world} ranks, 16-byte triplets from pinned host "
                        f"(xsb_insert_triplets), CSC slabs read back to host"}


def run_fd(args, xsb, xd, bench, rank, world, local, dev):
    """BASELINE.json configs[4]: fdrand 3-D n^3 (n = 400: 64 M unknowns, 767 M updateindex! calls, 447 M
    nnz), generation sharded by node slab, columns owned by the same slabs, STRONG scaling.  Only the
    (l, l + nx*ny) couplings that straddle a slab face change owner."""
    n1 = args.fd_n
    N = n1 ** 3
    splits = [N * r // world for r in range(world)] + [N]
    mode = xsb.DETERMINISTIC if args.mode == "deterministic" else xsb.FAST
    D = xd.DistExtendableSparseMatrix(N, N, splits=splits, device=local)
    h = D.h
    h.set_profiling(True)
    n_ins = xsb.capi.stream_count_fdrand(n1, n1, n1)

    def step():
        h.reset()
        h.emit_fdrand(n1, n1, n1, seed=20240717, flavour=xsb.UPDATE, l_range=(splits[rank], splits[rank + 1]))
        return D.flush(mode, wait=False)  # the 16-byte offsets all-gather is launched, not waited for

    for _ in range(args.warmup):
        step()
    h.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    h.timer_start()
    for _ in range(args.steps):
        nnz, _ = step()
    ms_local = h.timer_stop()
    torch.cuda.synchronize()
    dist.barrier()
    tmax = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    st = h.flush_stats()
    if rank == 0:
        peak, _ = bench.peaks()
        b_flush = bench.flush_bytes(n_ins, 0, int(D.nnz_global), N)
        line = {"metric": bench.METRIC, "value": n_ins / (ms_step / 1e3), "unit": bench.UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"fdrand 3-D {n1}^3 (BASELINE.json configs[4]), node-slab generation and column-slab "
                                       f"ownership over {world} ranks, updateindex! stream + flush!, CSC left sharded",
                           "mode": args.mode, "exchange": dict(D.last_exchange),
                           "host_phase_ms_last_step": [round(x, 3) for x in D.last_phase_ms]},
                "flush_rank0": {"ms": st["ms_total"], "column_path": st["column_path"], "n_inserted": st["n_inserted"],
                                "nnz_new": st["nnz_new"]},
                "whole_job_flush_fraction_of_hbm_peak": b_flush / (ms_step / 1e3) / 1e9 / (peak * world),
                "nnz_global": int(D.nnz_global), "n_inserted": int(n_ins)}
        print(json.dumps(line))
    h.close()
    dist.barrier()
    dist.destroy_process_group()
    import faulthandler

    faulthandler.cancel_dump_traceback_later()
