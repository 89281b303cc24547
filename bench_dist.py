"""N>1 arm of bench.py: weak scaling of the partitioned FEM assembly over column-owning ranks, plus
BASELINE.json configs[4] (fdrand 400^3, strong scaling) as the `cfg5` key of the same line.

Global mesh: 128 x 128 x (127*N + 1) nodes (Kuhn 6-tet).  Rank r emits the tetrahedra of its
127 cube layers (245 805 960 rawupdateindex! calls, the same per-GPU work as the N=1 workload)
and owns the columns of the z-planes [127 r, 127 (r+1)) (the last rank also owns the top plane).
Records whose column lies on the interface plane travel to the rank above in one grouped
NCCL send/receive between neighbours; every rank then merges into its own CSC slab.  The assembled matrix is left
sharded (slab-local colptr + global offset); that is what is timed.  The first (warm-up) step
counts what every rank sends; the following steps use the fixed-capacity exchange
(xsb_route_pack / xsb_route_unpack): no count visits the host, one grouped point-to-point launch per step.
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist


def bind_to_gpu_numa(local):
    """Pins this rank (and therefore the pinned host buffers it allocates afterwards: first touch) to the CPUs next
    to its GPU.  Returns a short description for the bench line."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpus = open(f"{base}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = ids & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"gpu": bdf, "numa_node": node, "cpus": cpus, "bound": bool(allowed)}
    except Exception as e:  # noqa: BLE001  (containers without sysfs: run unbound)
        return {"bound": False, "why": str(e)[:80]}


def timed_steps(h, step, steps, dev):
    """K steps on the device clock (CUDA events on the library's stream), barrier + synchronise on both sides, max over ranks."""
    h.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    h.timer_start()
    for _ in range(steps):
        step()
    ms_local = h.timer_stop()
    torch.cuda.synchronize()
    dist.barrier()
    tmax = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return float(tmax.item()) / steps


def run(args, xsb, rank, world, local):
    import bench

    # watchdog: a rank stuck in a collective dumps its Python stack and exits instead of hanging the box
    import faulthandler
    import sys

    faulthandler.dump_traceback_later(int(os.environ.get("XSB_BENCH_WATCHDOG_S", "400")), exit=True, file=sys.stderr)
    numa = bind_to_gpu_numa(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from xsparse_b200 import dist as xd

    dev = torch.device("cuda", local)
    if getattr(args, "workload", "fem") == "fd400":
        line = run_fd(args, xsb, xd, bench, rank, world, local, dev)
        if rank == 0:
            print(json.dumps(line))
        dist.barrier()
        dist.destroy_process_group()
        faulthandler.cancel_dump_traceback_later()
        return
    nx = ny = args.mesh
    layers = args.mesh - 1
    nz_nodes = layers * world + 1
    N = nx * ny * nz_nodes
    plane = nx * ny
    splits = [plane * layers * r for r in range(world)] + [N]
    mode = xsb.DETERMINISTIC if args.mode == "deterministic" else xsb.FAST
    D = xd.DistExtendableSparseMatrix(N, N, splits=splits, device=local)
    h = D.h
    n_ins_rank = xsb.capi.stream_count_p1fem(nx, ny, layers + 1)

    # Every rank visits its INTERFACE layer first (the cube layer under the plane the rank above owns: the only
    # elements with entries in foreign columns), so that those records can be handed to the exchange before the
    # interior is assembled (XSB_EXCHANGE_EARLY=1).  Same entries as the natural order; the stream order -- and
    # with it the reference result this is compared with (--verify) -- is [interface layer, interior layers] per rank.
    top = layers * (rank + 1) - 1

    def emit_interface():
        h.emit_p1fem(nx, ny, nz_nodes, flavour=xsb.RAW, cz_range=(top, top + 1))

    def emit_interior():
        h.emit_p1fem(nx, ny, nz_nodes, flavour=xsb.RAW, cz_range=(layers * rank, top))

    # 1: the interface blocks leave before the interior is assembled (DistExtendableSparseMatrix.exchange_begin).
    # Measured at N = 8 (B200, NVLink): 4.12 ms per step against 3.92 without -- the transfer is 31 MB per rank and
    # the ranks wait for each other ~0.05 ms either way, while the split costs launches: off by default.
    early = os.environ.get("XSB_EXCHANGE_EARLY", "0") == "1"

    def step():
        h.reset()
        emit_interface()
        if early:
            D.exchange_begin()
        emit_interior()
        return D.flush(mode, wait=False)  # the 16-byte offsets all-gather is launched, not waited for

    for _ in range(args.warmup):
        step()
    launches0 = h.kernel_launches
    with bench.ClockSampler(local) as clk:
        ms_step = timed_steps(h, step, args.steps, dev)
        launches_timed = h.kernel_launches - launches0
        # keep the clock sampler busy for about half a second: the SAME number of extra steps on every rank
        # (a step holds collectives; a per-rank time limit would let the ranks disagree and hang)
        n_fill = int(max(0.0, 500.0 - ms_step * args.steps) / max(ms_step, 1e-3)) + 1
        for _ in range(min(n_fill, 200)):
            step()
        torch.cuda.synchronize()
        dist.barrier()
    launches = torch.tensor([launches_timed], dtype=torch.int64, device=dev)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    value = world * n_ins_rank / (ms_step / 1e3)
    peak, peak_src = bench.peaks()
    # stage times of rank 0 (profiled steps: CUDA events around every stage of the flush; not the headline timing)
    h.set_profiling(True)
    stage, ms_emit = {}, 0.0
    prof_steps = min(args.steps, 3)
    st = None
    for _ in range(prof_steps):
        h.reset()
        h.timer_start()
        emit_interface()
        if early:
            D.exchange_begin()
        emit_interior()
        ms_emit += h.timer_stop()
        D.flush(mode, wait=False)
        st = h.flush_stats()
        for k, v in st.items():
            if k.startswith("ms_"):
                stage[k] = stage.get(k, 0.0) + v
    h.set_profiling(False)
    roof = bench.roofline(st, {k: v / prof_steps for k, v in stage.items()}, ms_emit / prof_steps, h.n, peak, peak_src, None)
    roof["kernel"] += " [rank 0]"
    exchange = dict(D.last_exchange)
    if D.last_device_phase_ms:  # XSB_DIST_TIMING=1 (diagnostic: the events synchronise every step)
        mine = {k: round(v, 4) for k, v in D.last_device_phase_ms.items()}
        mine["emit"] = round(ms_emit / prof_steps, 4)
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        exchange["device_phase_ms_last_step_per_rank"] = allp
    nnz_global = int(D.nnz_global)
    D_transport = D.transport_note or "nccl"
    parity = None
    if getattr(args, "verify", False):
        # the sharded result of the last step against the SAME global assembly on one GPU (rank 0 builds it with
        # a single handle): SHA-256 of every slab's colptr | rowval | nzval must be identical
        mine = bench.csc_digest(*h.fetch_csc_numpy())
        digests = [None] * world
        dist.all_gather_object(digests, mine)
        if rank == 0:
            import numpy as np

            g = xsb.Handle(N, N)
            for r in range(world):  # the rank-ordered concatenation of the ranks' streams
                t = layers * (r + 1) - 1
                g.emit_p1fem(nx, ny, nz_nodes, flavour=xsb.RAW, cz_range=(t, t + 1))
                g.emit_p1fem(nx, ny, nz_nodes, flavour=xsb.RAW, cz_range=(layers * r, t))
            g.flush(mode)
            cp, rv, nz = g.fetch_csc_numpy()
            g.close()
            want = []
            for r in range(world):
                lo, hi = splits[r], splits[r + 1]
                a, b = int(cp[lo]) - 1, int(cp[hi]) - 1
                want.append(bench.csc_digest(cp[lo:hi + 1] - cp[lo] + 1, rv[a:b], nz[a:b]))
            parity = {"slabs_identical_to_single_gpu_assembly": digests == want, "slab_sha256": digests}
            del cp, rv, nz
    D.close()
    del D

    cfg5 = None
    if not args.no_legs:
        cfg5 = run_fd(args, xsb, xd, bench, rank, world, local, dev)
    e2e = measure_e2e(args, xsb, xd, rank, world, local, mode)
    clocks = clk.summary()
    if rank == 0:
        cfg = bench.workload(args)
        cfg["workload"] = (f"P1-FEM Laplacian+mass, {nx}x{ny}x{nz_nodes}-node Kuhn mesh sharded over {world} ranks "
                           f"({layers} cube layers = {n_ins_rank} rawupdateindex! calls per rank), column-slab ownership, "
                           f"interface layer assembled first; its records travel to the neighbouring slab as one fixed-capacity "
                           f"block per step ({D_transport}"
                           f"{', sent before the interior is assembled' if early else ''}), CSC left sharded")
        cfg["parallelism"] = f"column-slab x{world}"
        cfg["exchange"] = exchange
        cfg["numa"] = numa
        line = {
            "metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": roof,
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches.item()), "clocks": clocks,
            "nnz_global": nnz_global, "n_inserted": int(world * n_ins_rank), "cfg5": cfg5,
        }
        if parity is not None:
            line["parity_checked"] = parity
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


def measure_e2e(args, xsb, xd, rank, world, local, mode):
    """End to end at N ranks with HOST buffers: every step each rank copies its insertion stream (16-byte
    triplets) from pinned host memory (allocated after the rank was bound to its GPU's NUMA node), inserts,
    routes, flushes and reads its CSC slab back to pinned host memory."""
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
    ms = timed_steps(g, step, steps, torch.device("cuda", local))
    tot = torch.tensor([cnt, 16 * cnt, 8 * (g.n + 1) + 16 * int(nnz)], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    D.close()
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
    n_ins = xsb.capi.stream_count_fdrand(n1, n1, n1)

    # interface nodes first: the first and the last z-plane of the rank's node slab hold the only nodes with
    # neighbours (= columns) in other slabs
    plane = n1 * n1
    lo, hi = splits[rank], splits[rank + 1]
    parts = [(lo, min(lo + plane, hi)), (max(hi - plane, min(lo + plane, hi)), hi)]
    interior = (parts[0][1], parts[1][0])

    def step():
        h.reset()
        for a, b in parts:
            if b > a:
                h.emit_fdrand(n1, n1, n1, seed=20240717, flavour=xsb.UPDATE, l_range=(a, b))
        if os.environ.get("XSB_EXCHANGE_EARLY", "0") == "1":
            D.exchange_begin()
        if interior[1] > interior[0]:
            h.emit_fdrand(n1, n1, n1, seed=20240717, flavour=xsb.UPDATE, l_range=interior)
        return D.flush(mode, wait=False)  # the 16-byte offsets all-gather is launched, not waited for

    for _ in range(max(2, min(args.warmup, 3))):
        step()
    steps = max(1, min(args.steps, 5))
    ms_step = timed_steps(h, step, steps, dev)
    h.set_profiling(True)
    step()
    st = h.flush_stats()
    h.set_profiling(False)
    nnz_global = int(D.nnz_global)
    peak, _ = bench.peaks()
    b_flush = bench.flush_bytes(n_ins, 0, nnz_global, N)
    line = {"metric": bench.METRIC, "value": n_ins / (ms_step / 1e3), "entries_per_s": n_ins / (ms_step / 1e3),
            "unit": bench.UNIT, "n_gpus": world,
            "steps": steps, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"fdrand 3-D {n1}^3 (BASELINE.json configs[4]), node-slab generation and column-slab "
                                   f"ownership over {world} ranks, updateindex! stream + flush!, CSC left sharded",
                       "mode": args.mode, "exchange": dict(D.last_exchange), "fixed_capacity_steps": D.fixed_steps},
            "flush_rank0": {"ms": st["ms_total"], "column_path": st["column_path"], "n_inserted": st["n_inserted"],
                            "nnz_new": st["nnz_new"]},
            "whole_job_flush_fraction_of_hbm_peak": b_flush / (ms_step / 1e3) / 1e9 / (peak * world),
            "nnz_global": nnz_global, "n_inserted": int(n_ins)}
    D.close()
    return line
