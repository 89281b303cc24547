"""world_size-2/3 gloo tests (CPU) of the multi-rank choreography in dist.py: owner bucketing,
count exchange, all-to-all-v of the records, slab merge order, global colptr offsets.
The per-rank slab object is a TEST DOUBLE built on the CPU oracle (the product backend is the
CUDA library and needs a GPU); what is under test here is the host-side exchange logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleSlabBackend:
    """Stands in for capi.Handle(slab=...) on CPU tensors."""

    def __init__(self, ora, m, n, world, rank, splits):
        self.ora, self.m, self.n_global = ora, m, n
        self.world, self.rank, self.splits = world, rank, list(splits)
        self.col_begin, self.width = splits[rank], splits[rank + 1] - splits[rank]
        self.A = ora.OracleExt(m, self.width)
        self.staged = []
        self.routed = None

    def insert_batch(self, I, J, V, flavour=0):
        for i, j, v in zip(I, J, V):
            assert 1 <= i <= self.m and 1 <= j <= self.n_global
            self.staged.append((int(i), int(j), float(v), int(flavour)))

    @property
    def pending(self):
        return len(self.staged)

    def _owners(self):
        return np.searchsorted(np.asarray(self.splits[1:]), [j - 1 for (_, j, _, _) in self.staged], side="right")

    def route_count(self):
        return np.bincount(self._owners(), minlength=self.world).tolist()

    def route_prepare(self, send, capacity):
        owners = self._owners()
        counts = np.bincount(owners, minlength=self.world).tolist()
        assert capacity >= len(self.staged) - counts[self.rank]
        buf = send.numpy()
        k = 0
        for d in range(self.world):  # off-rank buckets only, destination after destination, stream order
            if d == self.rank:
                continue
            for s in np.nonzero(owners == d)[0]:
                i, j, v, fl = self.staged[s]
                buf[2 * k] = ((j - 1) * self.m + (i - 1)) * 4 + fl
                buf[2 * k + 1] = np.float64(v).view(np.int64)
                k += 1
        self.own = [(i, j - self.col_begin, v, fl) for (i, j, v, fl), o in zip(self.staged, owners) if o == self.rank]
        self.low, self.high = [], []
        self.staged = []
        self.routed = True
        return counts

    def route_finish(self, src, recv, count):
        assert src != self.rank
        buf = recv.numpy()
        out = self.low if src < self.rank else self.high
        assert src < self.rank or True
        for k in range(count):
            key = int(buf[2 * k])
            fl, ij = key % 4, key // 4
            i, j = ij % self.m + 1, ij // self.m + 1
            assert self.col_begin < j <= self.col_begin + self.width, "record routed to the wrong owner"
            out.append((i, j - self.col_begin, float(np.int64(buf[2 * k + 1]).view(np.float64)), fl))

    def flush(self, mode=0):
        before = self.A.nnz
        for (i, j, v, fl) in self.low + self.own + self.high:  # lower ranks, own, higher ranks
            if fl == 0:
                self.A.updateindex(v, i, j)
            elif fl == 1:
                self.A.rawupdateindex(v, i, j)
            else:
                self.A[i, j] = v
        self.routed = None
        nnz = self.A.nnz
        return nnz, nnz != before

    def synchronize(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, n, splits, out_q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as ora
        import importlib.util

        spec = importlib.util.spec_from_file_location("xsb_dist", os.path.join(root, "extendablesparse.jl_b200", "dist.py"))
        xd = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(xd)

        backend = OracleSlabBackend(ora, m, n, world, rank, splits)
        D = xd.DistExtendableSparseMatrix(m, n, splits=splits, backend=backend)
        rng = np.random.default_rng(100 + rank)
        streams = []
        results = []
        for splice in range(2):
            cnt = 400 + 50 * rank
            I = rng.integers(1, m + 1, cnt)
            J = rng.integers(1, n + 1, cnt)
            V = rng.standard_normal(cnt)
            V[::9] = 0.0
            fl = splice  # update flavour, then raw
            D.insert_batch(I, J, V, fl)
            streams.append((I, J, V, fl))
            D.exchange_begin()  # without the peer transport (this backend; counted steps): nothing to do
            if splice % 2 == 0:
                nnz, changed = D.flush()
            else:  # offsets all-gather launched only; the global values are read on demand
                nnz, _local = D.flush(wait=False)
                changed = D.changed_any
            cp, rv, nz = backend.A.csc()
            results.append((nnz, changed, D.nnz_offset, D.nnz_global, cp, rv, nz, dict(D.last_exchange)))
        assert D.transport_note == "" and D._peer is None  # the CPU double never leaves the counted exchange
        D.recount()  # collective no-op here; the next flush counts (as every flush of this backend)
        out_q.put((rank, streams, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,splits", [(2, None), (3, [0, 5, 6, 40]), (2, [0, 39, 40])])
def test_distributed_assembly_matches_serial(oracle, world, splits):
    m, n = 30, 40
    if splits is None:
        splits = [0, 20, 40]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, splits, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, streams, results = q.get(timeout=120)
        got[rank] = (streams, results)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # serial reference: the ranks' streams concatenated in rank order, splice by splice
    A = oracle.OracleExt(m, n)
    for splice in range(2):
        for r in range(world):
            I, J, V, fl = got[r][0][splice]
            A.insert_batch(I, J, V, fl)
        cp, rv, nz = A.csc()
        off = 0
        sent = recv = 0
        for r in range(world):
            nnz, changed, nnz_offset, nnz_global, lcp, lrv, lnz, ex = got[r][1][splice]
            lo, hi = splits[r], splits[r + 1]
            assert nnz_offset == off and nnz_global == len(nz)
            # slab colptr shifted by the offset == the global colptr restricted to the slab
            assert np.array_equal(lcp + nnz_offset, cp[lo:hi + 1])
            assert np.array_equal(lrv, rv[cp[lo] - 1:cp[hi] - 1])
            assert np.array_equal(lnz.view(np.uint64), nz[cp[lo] - 1:cp[hi] - 1].view(np.uint64))
            assert changed is True
            off += nnz
            sent += ex["sent_off_rank"]
            recv += ex["received_off_rank"]
            kept = ex["kept"] + (0 if r else 0)
            assert kept + ex["sent_off_rank"] == len(got[r][0][splice][2])
        assert off == len(nz)
        total = sum(len(got[r][0][splice][2]) for r in range(world))
        assert recv == sent and 0 < sent < total  # the own bucket never travels
