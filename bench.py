#!/usr/bin/env python
"""bench.py -- assembled entries/s of insert + flush! on B200, with roofline and CPU baseline.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            our arm (libxsparse_b200 through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU algorithm (oracle port)

A step = one full assembly of the workload: reset!, insertion of the whole (i,j,v) stream,
flush! into a fresh CSC.  N=1 workload: BASELINE.json configs[1], the 3-D P1-FEM Laplacian on
a 128^3-node Kuhn mesh (245 805 960 rawupdateindex! calls, 31 065 598 nnz, Float64/Int64).
N>1: weak scaling -- every rank assembles its own 128x128x128-node slab of a 128x128x(127N+1)
mesh; columns are owned by z-slab and the interface plane is routed with an NCCL all-to-all.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "assembled entries/sec (insert+flush!)"
UNIT = "entries/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(s[2]) for s in self.samples)}


# ---------------------------------------------------------------------------------- workloads
def workload(args):
    n = args.mesh
    return {"workload": f"P1-FEM Laplacian+mass, {n}^3-node Kuhn 6-tet mesh, rawupdateindex! stream + flush! "
                        f"(BASELINE.json configs[1])",
            "nodes": n ** 3, "mesh": n, "Tv": "Float64", "Ti": "Int64", "mode": args.mode,
            "l2_policy": "inputs (3.9 GB of records) far exceed the 126 MB L2; no explicit flush"}


def flush_bytes(n_ins, nnz_old, nnz_new, ncols):
    """Algorithmic bytes of one flush!, SURVEY.md 8(d): 16 B per triplet (packed key + Float64),
    16 B per CSC entry (Int64 row + Float64), 8 B per colptr entry; old CSC read, new CSC written."""
    return 16 * n_ins + 16 * nnz_old + 8 * (ncols + 1) + 16 * nnz_new + 8 * (ncols + 1)



# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of the
# same workload (profiles/r1e_ncu_full_fem128.csv, FEM 128^3); other workloads / kernels: null
NCU_TRAFFIC_FEM128 = {"group_scatter_kernel": 4.0631e9 + 3.8998e9, "colthread_direct_kernel": 4.1566e9 + 0.4968e9,
                      "group_count_kernel": 0.9834e9 + 0.2440e9}


def roofline(st, ms, ncols, peak, peak_src, traffic):
    """Roofline of the DOMINANT kernel of the flush (largest device time per step, CUDA events on the
    library's stream around that kernel) + the whole-flush fraction on SURVEY.md 8(d)'s bytes.
    Algorithmic bytes per launch: records are 16 B, CSC entries 16 B."""
    rec = st["n_inserted"] + st["nnz_old"]
    nnz = st["nnz_new"]
    pairs = st["group_pairs"]
    path = st["column_path"]
    cands = {}
    if path == 3:
        # counting at insertion: a share `pre` of the records is counted from the 4-byte column ids their producer
        # left (or was counted by the producer itself); a pair costs 16 (pair record) + 4 (column) + 2 (count) bytes
        pre = float(st.get("precounted", 0.0))
        cands["group_count_kernel (+ pair_totals: column ids / records in, (column, chunk) pairs out)"] = (
            ms["ms_group_count"], (16 - 12 * pre) * rec + 22 * pairs, 1)
        cands["group_scatter_kernel (stable scatter of every record to its column)"] = (
            ms["ms_group_scatter"], 32 * rec + 8 * pairs, 1)
        if st["sort_passes"]:
            cands["onesweep_kernel on the (column, chunk) pairs (+ pair offsets)"] = (
                ms["ms_pair_sort"], 32 * pairs * st["sort_passes"] + 24 * pairs, st["sort_passes"])
        else:
            cands["pair buckets (colpair scans + pair_bucket_kernel + pair_offsets_kernel)"] = (
                ms["ms_pair_sort"], 12 * ncols + 26 * pairs, 1)
    else:
        cands["onesweep_kernel (one radix pass over 16-B records)"] = (ms["ms_sort"], 32 * rec * st["sort_passes"],
                                                                      st["sort_passes"])
    if path >= 2:
        cands["colthread_direct_kernel (thread-per-column fold writing rowval / nzval / colptr)" if st.get("direct_fold")
              else "colthread_kernel (+ leftover colfold_kernel: per-column fold, read records, park entries)"] = (
            ms["ms_fold"], 16 * rec + 16 * nnz, 1)
        cands["compact_entries_kernel (parked entries -> rowval / nzval)"] = (ms["ms_compact"], 32 * nnz, 1)
    else:
        cands["reduce_emit_kernel"] = (ms["ms_reduce"], 16 * rec + 16 * nnz, 1)
    name = max(cands, key=lambda k: cands[k][0])
    t, byts, launches = cands[name]
    achieved = byts / (t / 1e3) / 1e9
    if traffic is None and st["n_inserted"] == 245805960 and st["nnz_old"] == 0:
        traffic = next((v for k, v in NCU_TRAFFIC_FEM128.items() if name.startswith(k)), None)
    b_flush = flush_bytes(st["n_inserted"], st["nnz_old"], nnz, ncols)
    flush_gbs = b_flush / (ms["ms_total"] / 1e3) / 1e9
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "bytes_per_launch": byts / launches, "ms_per_launch": t / launches, "launches_per_step": launches,
            "kernels": {k: {"ms": v[0], "GB/s": v[1] / (v[0] / 1e3) / 1e9 if v[0] > 0 else None,
                            "frac": v[1] / (v[0] / 1e3) / 1e9 / peak if v[0] > 0 else None} for k, v in cands.items()},
            "flush": {"algorithmic_bytes": b_flush, "ms": ms["ms_total"], "achieved": flush_gbs,
                      "frac": flush_gbs / peak, "column_path": path},
            "stage_ms_per_step": dict(sorted(ms.items()))}


# ---------------------------------------------------------------------------------- CPU legs
def cpu_reference_leg(mesh, steps, warmup):
    """Times the oracle port of the reference's serial path (ExtendableSparseMatrix +
    rawupdateindex! + flush!) on one host core; returns entries/s (best step) and seconds/step."""
    from oracle import oracle as ora

    ora.build()
    I, J, V = ora.fem_stream(mesh, mesh, mesh)
    n = mesh ** 3
    times = []
    for it in range(warmup + steps):
        A = ora.OracleExt(n, n)
        t0 = time.perf_counter()
        A.insert_batch(I, J, V, ora.RAW)
        A.flush()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        del A
    return len(V), times


def host_threads():
    """Host threads the CPU legs may use: the cores this process is allowed on, capped at 64
    (every partition of the MT path owns a dense column map)."""
    try:
        return max(1, min(64, len(os.sched_getaffinity(0))))
    except AttributeError:
        return max(1, min(64, os.cpu_count() or 1))


def cpu_reference_mt_leg(mesh, steps, warmup, nthreads):
    """Times the oracle port of the reference's PARTITIONED path (MTExtendableSparseMatrixCSC =
    GenericMTExtendableSparseMatrixCSC{SparseMatrixDILNKC}; loop of test/femtools.jl:75-110):
    2*nthreads slab partitions in two colours inserted by `nthreads` host threads, then the
    reference's serial flush! (Base.sum(xmatrices, csc), sparsematrixdilnkc.jl:397-435)."""
    import numpy as np

    from oracle import oracle as ora

    ora.build()
    I, J, V = ora.fem_stream(mesh, mesh, mesh)
    n = mesh ** 3
    nparts = 2 * nthreads
    ncells = len(V) // 20
    pb = 20 * (np.arange(nparts + 1, dtype=np.int64) * ncells // nparts)
    times = []
    for it in range(warmup + steps):
        A = ora.OracleMT(n, n, nparts)
        t0 = time.perf_counter()
        A.insert_partitioned(I, J, V, pb, nthreads, ora.RAW)
        A.flush()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        del A
    return len(V), times


def cpu_baseline_both(mesh, steps, warmup):
    """Serial and partitioned CPU legs on the same sample; the faster one is the baseline."""
    nth = host_threads()
    n_ins, ts = cpu_reference_leg(mesh, steps, warmup)
    serial = n_ins * len(ts) / sum(ts)
    mt = None
    if nth > 1:
        _, tm = cpu_reference_mt_leg(mesh, steps, warmup, nth)
        mt = n_ins * len(tm) / sum(tm)
    use_mt = mt is not None and mt > serial
    times = tm if use_mt else ts
    sample = (f"P1-FEM {mesh}^3-node Kuhn mesh ({n_ins} insertions per assembly), same generator and flavour as "
              f"the GPU workload; C port (gcc -O3) of the reference's algorithm. Serial ExtendableSparseMatrix "
              f"(rawupdateindex! + flush!, 1 thread): {serial / 1e9:.4f} G entries/s. "
              + (f"Partitioned MTExtendableSparseMatrixCSC ({2 * nth} slab partitions, 2 colours, {nth} threads "
                 f"inserting, serial flush! as in the reference): {mt / 1e9:.4f} G entries/s. "
                 if mt is not None else "One host core available: partitioned path not timed. ")
              + f"Reported: the {'partitioned' if use_mt else 'serial'} path")
    cpu = {"value": max(serial, mt or 0.0), "unit": UNIT, "cores": nth if use_mt else 1, "kind": "port",
           "sample": sample, "serial_value": serial, "mt_value": mt, "mt_threads": nth if mt is not None else None}
    return n_ins, times, cpu


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh = args.ref_mesh
    n_ins, times, cpu = cpu_baseline_both(mesh, args.steps, args.warmup)
    total = sum(times)
    value = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch

    import __graft_entry__ as ge

    ge.build()
    import xsparse_b200 as xsb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus != 1 and world == 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        import bench_dist

        return bench_dist.run(args, xsb, rank, world, local)

    mesh = args.mesh
    n = mesh ** 3
    mode = xsb.DETERMINISTIC if args.mode == "deterministic" else xsb.FAST
    h = xsb.Handle(n, n, device=local)
    h.set_profiling(True)
    n_ins = xsb.capi.stream_count_p1fem(mesh, mesh, mesh)

    def step():
        h.reset()
        h.emit_p1fem(mesh, mesh, mesh, flavour=xsb.RAW)
        return h.flush(mode)

    for _ in range(args.warmup):
        step()
    h.synchronize()
    torch.cuda.synchronize()
    launches0 = h.kernel_launches
    stage = {}
    with ClockSampler(local) as clk:
        h.timer_start()
        for _ in range(args.steps):
            nnz, _ = step()
            st = h.flush_stats()
            for k, v in st.items():
                if k.startswith("ms_"):
                    stage[k] = stage.get(k, 0.0) + v
        ms = h.timer_stop()
        launches_timed = h.kernel_launches - launches0
        # keep the sampler alive for at least a few samples on very short runs
        t_end = time.time() + max(0.0, 0.5 - ms / 1e3)
        while time.time() < t_end:
            step()
    st = h.flush_stats()
    ms_step = ms / args.steps
    value = n_ins / (ms_step / 1e3)
    peak, peak_src = peaks()

    roof = roofline(st, {k: v / args.steps for k, v in stage.items()}, n, peak, peak_src, args.traffic)

    # ---- end-to-end through the C ABI with HOST buffers
    e2e = measure_e2e(args, xsb, h, mesh, n, n_ins, mode, np, torch)

    # ---- CPU baseline (rank 0, N=1): the oracle port on the same workload
    cpu = None
    if not args.no_cpu:
        _, _, cpu = cpu_baseline_both(args.cpu_mesh, 1, 0)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload(args),
        "roofline": roof,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_timed),
        "clocks": clk.summary(), "nnz": int(nnz), "n_inserted": int(n_ins),
    }
    print(json.dumps(line))
    h.close()


def measure_e2e(args, xsb, h, mesh, n, n_ins, mode, np, torch):
    """Same metric through the public C ABI with HOST buffers.  Every step copies the insertion stream
    host->device from pinned memory, runs insert + flush!, and reads the CSC (colptr, rowval, nzval)
    back to pinned host memory.  Headline: the stream as 16-byte triplets (xsb_insert_triplets, the
    buffer the glue's updateindex! appends to: 16 B per insertion over PCIe); beside it the same
    stream as three Int64/Int64/Float64 arrays (xsb_insert_batch: 24 B per insertion)."""
    emesh = args.e2e_mesh
    en = emesh ** 3
    g = xsb.Handle(en, en, device=h.device)
    g.set_precount(False)  # the host arrays hold the stream in call order, as user code would produce it
    g.emit_p1fem(emesh, emesh, emesh, flavour=xsb.RAW)
    cnt = g.pending
    dI = torch.empty(cnt, dtype=torch.int64, device="cuda")
    dJ = torch.empty(cnt, dtype=torch.int64, device="cuda")
    dV = torch.empty(cnt, dtype=torch.float64, device="cuda")
    c = xsb.capi
    import ctypes as C

    got = C.c_int64(0)
    c.check(c.lib().xsb_debug_fetch_staged(g._h, 0, dI.data_ptr(), dJ.data_ptr(), dV.data_ptr(), None, cnt,
                                           C.byref(got)), g._h)
    hI = torch.empty(cnt, dtype=torch.int64, pin_memory=True)
    hJ = torch.empty(cnt, dtype=torch.int64, pin_memory=True)
    hV = torch.empty(cnt, dtype=torch.float64, pin_memory=True)
    hI.copy_(dI)
    hJ.copy_(dJ)
    hV.copy_(dV)
    # the same stream as xsb_triplet {u32 row, u32 col, f64 val}: two 64-bit words per insertion
    dT = torch.empty((cnt, 2), dtype=torch.int64, device="cuda")
    dT[:, 0] = dI | (dJ << 32)
    dT[:, 1] = dV.view(torch.int64)
    hT = torch.empty((cnt, 2), dtype=torch.int64, pin_memory=True)
    hT.copy_(dT)
    torch.cuda.synchronize()
    del dI, dJ, dV, dT
    g.reset()
    g.set_precount(True)
    # result buffers (pinned), sized after one dry run
    g.insert_triplets(hT, xsb.RAW, 0, cnt)
    nnz, _ = g.flush(mode)
    ocp = torch.empty(en + 1, dtype=torch.int64, pin_memory=True)
    orv = torch.empty(nnz, dtype=torch.int64, pin_memory=True)
    onz = torch.empty(nnz, dtype=torch.float64, pin_memory=True)

    def step_triplets():
        g.reset()
        g.insert_triplets(hT, xsb.RAW, 0, cnt)
        g.flush(mode)
        g.fetch_csc(ocp, orv, onz)

    def step_ijv():
        g.reset()
        g.insert_batch(hI, hJ, hV, xsb.RAW, count=cnt)
        g.flush(mode)
        g.fetch_csc(ocp, orv, onz)

    def timed(step):
        for _ in range(max(1, min(args.warmup, 2))):
            step()
        g.synchronize()
        steps = max(1, min(args.steps, 3))
        g.timer_start()
        for _ in range(steps):
            step()
        return g.timer_stop() / steps

    ms_ijv = timed(step_ijv)
    check_a = (int(orv[:1000].sum()), float(onz[:1000].sum()))
    ms = timed(step_triplets)
    assert check_a == (int(orv[:1000].sum()), float(onz[:1000].sum())), "triplet and (I,J,V) assemblies differ"
    g.close()
    d2h = 8 * (en + 1) + 16 * int(nnz)
    return {"value": cnt / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": 16 * cnt,
            "d2h_bytes_per_step": d2h, "ms_per_step": ms,
            "workload": f"P1-FEM {emesh}^3-node mesh, {cnt} insertions from a pinned host array of 16-byte triplets "
                        f"(xsb_insert_triplets), CSC read back to host",
            "ijv": {"value": cnt / (ms_ijv / 1e3), "unit": UNIT, "h2d_bytes_per_step": 24 * cnt,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_ijv,
                    "workload": "same stream as three pinned host arrays Int64 I, Int64 J, Float64 V "
                                "(xsb_insert_batch)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh", type=int, default=128, help="nodes per direction of the FEM mesh (128 = configs[1])")
    ap.add_argument("--e2e-mesh", type=int, default=128)
    ap.add_argument("--cpu-mesh", type=int, default=128, help="mesh of the cpu_baseline sample")
    ap.add_argument("--ref-mesh", type=int, default=128, help="mesh of one --impl reference step (128 = the GPU arm's workload)")
    ap.add_argument("--mode", default="deterministic", choices=["deterministic", "fast"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="fem", choices=["fem", "fd400"],
                    help="N>1 only: fem = weak scaling of configs[1] (default, the bench line); fd400 = configs[4], "
                         "fdrand 400^3 sharded by node slab over the ranks (strong scaling, no e2e leg)")
    ap.add_argument("--fd-n", type=int, default=400, help="grid points per direction of --workload fd400")
    ap.add_argument("--traffic", type=float, default=None,
                    help="ncu dram bytes per launch of the dominant kernel (default: the committed capture's figure)")
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    args.warmup = max(3, args.warmup) if args.impl == "ours" else max(0, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
