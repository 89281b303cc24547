"""Python mirror of the Julia drop-in glue (julia/ExtendableSparseB200.jl), call for call.

The reference plugs an insert buffer into its wrapper matrices through
`AbstractSparseMatrixExtension` (src/matrix/abstractsparsematrixextension.jl:1-19).  This module
restates BOTH sides of that boundary so that the sequence of C-ABI calls the Julia glue makes can be
executed (and timed) without a Julia runtime:

    SparseMatrixB200                      the extension: per-partition HOST buffer of 16-byte triplets
    sum_extensions(exts, csc)             Base.sum(exts, csc): xsb_set_csc -> xsb_insert_triplets per
                                          partition and flavour run -> xsb_flush -> xsb_fetch_csc, ONE handle
    GenericMTExtendableSparseMatrixCSC    the reference's wrapper (genericmtextendablesparsematrixcsc.jl:1-114):
                                          CSC hits are updated in place on the host, misses go to xmatrices[tid]

Indices are 1-based like the reference's.  No CPU fallback: the flush needs libxsparse_b200 and a GPU.
"""
from __future__ import annotations

import operator
from typing import List, Sequence, Tuple

import numpy as np

from . import capi

_CHUNK = 1 << 16
Csc = Tuple[np.ndarray, np.ndarray, np.ndarray]  # (colptr, rowval, nzval), 1-based int64 / float64


class SparseMatrixB200:
    """T_ext(m, n): the calls of one partition since the last flush!, in call order, in host memory."""

    def __init__(self, m: int, n: int, buffer: np.ndarray | None = None):
        self.m, self.n = int(m), int(n)
        if self.m >= 2 ** 32 or self.n >= 2 ** 32:
            raise ValueError("triplet buffers carry 32-bit indices")
        self.T = np.empty(_CHUNK, capi.TRIPLET_DTYPE) if buffer is None else buffer
        self.fill = 0
        self.runs: List[Tuple[int, int]] = []  # (index of the first triplet, flavour) of every run

    @property
    def nnz(self) -> int:
        """Upper bound of the distinct new entries (SparseArrays.nnz of the contract)."""
        return self.fill

    def size(self):
        return (self.m, self.n)

    def _grow(self, need: int):
        if need > len(self.T):
            cap = len(self.T)
            while cap < need:
                cap *= 2
            T = np.empty(cap, capi.TRIPLET_DTYPE)
            T[: self.fill] = self.T[: self.fill]
            self.T = T

    def push(self, flavour: int, v: float, i: int, j: int):
        if not (1 <= i <= self.m and 1 <= j <= self.n):
            raise capi.XsbBoundsError(capi.EBOUNDS, f"BoundsError: attempt to access {self.m}x{self.n} matrix at [{i}, {j}]")
        self._grow(self.fill + 1)
        if not self.runs or self.runs[-1][1] != flavour:
            self.runs.append((self.fill, flavour))
        self.T[self.fill] = (i, j, v)
        self.fill += 1

    def push_batch(self, flavour: int, I, J, V):
        """k-th element == k-th call."""
        I, J, V = np.asarray(I), np.asarray(J), np.asarray(V, np.float64)
        cnt = len(V)
        if cnt == 0:
            return
        if I.min() < 1 or I.max() > self.m or J.min() < 1 or J.max() > self.n:
            raise capi.XsbBoundsError(capi.EBOUNDS, "BoundsError: batch holds a position outside the matrix")
        self._grow(self.fill + cnt)
        if not self.runs or self.runs[-1][1] != flavour:
            self.runs.append((self.fill, flavour))
        t = self.T[self.fill: self.fill + cnt]
        t["row"], t["col"], t["val"] = I, J, V
        self.fill += cnt

    # the calls the generic wrappers forward to (genericmt...:87-114, genericextendable...:44-92)
    def rawupdateindex(self, op, v, i, j, tid=None):
        self.push(capi.RAW, _signed(op, v), i, j)

    def updateindex(self, op, v, i, j):
        self.push(capi.UPDATE, _signed(op, v), i, j)

    def __setitem__(self, ij, v):
        self.push(capi.ASSIGN, v, ij[0], ij[1])


def _signed(op, v):
    if op is operator.add or op == "+":
        return v
    if op is operator.sub or op == "-":
        return -v
    raise NotImplementedError("libxsparse_b200 implements op in {+,-}; other ops stay on the CPU buffers")


# One device handle per (size, partitions): created at the first flush!, reused by the following ones.
_HANDLES: dict = {}


def device_handle(m: int, n: int, nparts: int, device: int = 0) -> capi.Handle:
    key = (int(m), int(n), int(nparts), int(device))
    h = _HANDLES.get(key)
    if h is None or not h._h:
        h = capi.Handle(int(m), int(n), capi.I64, 1, int(nparts), device)
        _HANDLES[key] = h
    return h


def release_handles():
    for h in _HANDLES.values():
        h.close()
    _HANDLES.clear()


def sum_extensions(exts: Sequence[SparseMatrixB200], m: int, n: int, csc: Csc, mode=capi.DETERMINISTIC, device=0,
                   out: Csc | None = None, handle: capi.Handle | None = None) -> Tuple[Csc, bool]:
    """Base.sum(exts, csc): the flush of the plug-in contract.  Returns (new csc, pattern changed).

    `out` = caller-allocated (colptr, rowval, nzval) to fetch into (rowval / nzval with room for the new
    entries; e.g. pinned buffers), else fresh numpy arrays are allocated after xsb_flush returned nnz."""
    if sum(e.nnz for e in exts) == 0:
        return csc, False
    h = handle if handle is not None else device_handle(m, n, len(exts), device)
    colptr, rowval, nzval = csc
    h.set_csc(colptr, rowval if len(rowval) else None, nzval if len(nzval) else None)  # old CSC -> HBM
    for t, e in enumerate(exts):
        for r, (first, flavour) in enumerate(e.runs):
            last = e.runs[r + 1][0] if r + 1 < len(e.runs) else e.fill
            h.insert_triplets(e.T[first:last], flavour, t, last - first)
    nnz, changed = h.flush(mode)
    if out is None:
        out = (np.empty(n + 1, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float64))
    h.fetch_csc(out[0], out[1] if nnz else None, out[2] if nnz else None)
    return (out[0], out[1][:nnz], out[2][:nnz]), changed


class GenericMTExtendableSparseMatrixCSC:
    """GenericMTExtendableSparseMatrixCSC{SparseMatrixB200} (genericmtextendablesparsematrixcsc.jl:1-114)."""

    def __init__(self, m: int, n: int | None = None, nparts: int = 1, mode=capi.DETERMINISTIC, device: int = 0):
        self.m = int(m)
        self.n = int(m if n is None else n)
        self.nparts = int(nparts)
        self.mode, self.device = mode, device
        self.cscmatrix: Csc = (np.ones(self.n + 1, np.int64), np.empty(0, np.int64), np.empty(0, np.float64))
        self.xmatrices = [SparseMatrixB200(self.m, self.n) for _ in range(self.nparts)]
        self.pattern_changes = 0
        self._keys = np.empty(0, np.int64)  # (col-1)*m + (row-1) of the CSC entries: sorted

    # findindex(csc, i, j): sparsematrixcsc.jl:7-23 (0 = absent; here -1)
    def _find(self, I, J):
        I, J = np.asarray(I, np.int64), np.asarray(J, np.int64)
        if I.size and (I.min() < 1 or I.max() > self.m or J.min() < 1 or J.max() > self.n):
            raise capi.XsbBoundsError(capi.EBOUNDS, "BoundsError: position outside the matrix")
        q = (J - 1) * self.m + (I - 1)
        k = np.searchsorted(self._keys, q)
        k = np.minimum(k, max(len(self._keys) - 1, 0))
        hit = (self._keys[k] == q) if len(self._keys) else np.zeros(q.shape, bool)
        return np.where(hit, k, -1)

    def _update(self, flavour, op, v, i, j, tid):
        k = int(self._find([i], [j])[0])
        if k >= 0:  # CSC hit: in place (genericmt...:93-95,108-110)
            nz = self.cscmatrix[2]
            nz[k] = nz[k] + _signed(op, v)
        else:
            self.xmatrices[tid - 1].push(flavour, _signed(op, v), i, j)

    def rawupdateindex(self, op, v, i, j, tid=1):
        self._update(capi.RAW, op, v, i, j, tid)

    def updateindex(self, op, v, i, j, tid=1):
        self._update(capi.UPDATE, op, v, i, j, tid)

    def update_batch(self, flavour, I, J, V, tid=1):
        """The same calls in bulk (k-th element == k-th call): hits folded in place in call order, misses staged."""
        I, J, V = np.asarray(I, np.int64), np.asarray(J, np.int64), np.asarray(V, np.float64)
        k = self._find(I, J)
        hit = k >= 0
        if hit.any():
            np.add.at(self.cscmatrix[2], k[hit], V[hit])  # sequential in call order: the reference's left fold
        if (~hit).any():
            self.xmatrices[tid - 1].push_batch(flavour, I[~hit], J[~hit], V[~hit])

    def __setitem__(self, ij, v):
        k = int(self._find([ij[0]], [ij[1]])[0])
        if k < 0:  # genericmt...:63-68
            raise capi.XsbIllegalError(capi.EILLEGAL, "use rawupdateindex! for new entries into GenericMTExtendableSparseMatrixCSC")
        self.cscmatrix[2][k] = v

    def __getitem__(self, ij):
        k = int(self._find([ij[0]], [ij[1]])[0])
        if k >= 0:
            return float(self.cscmatrix[2][k])
        if self.nnznew == 0:
            return 0.0
        raise capi.XsbIllegalError(capi.EILLEGAL, "flush! GenericMTExtendableSparseMatrixCSC before using getindex")  # :71-82

    @property
    def nnznew(self) -> int:
        return sum(x.nnz for x in self.xmatrices)

    def flush(self):
        """flush!(ext): genericmt...:45-51 = Base.sum(xmatrices, csc), then fresh buffers."""
        if self.nnznew > 0:
            self.cscmatrix, changed = sum_extensions(self.xmatrices, self.m, self.n, self.cscmatrix, self.mode, self.device)
            self.pattern_changes += int(changed)
            self.xmatrices = [SparseMatrixB200(self.m, self.n) for _ in range(self.nparts)]
            cp, rv, _ = self.cscmatrix
            cols = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(cp))
            self._keys = cols * self.m + (rv - 1)
        return self

    def sparse(self) -> Csc:
        self.flush()
        return self.cscmatrix

    def reset(self):
        self.__init__(self.m, self.n, self.nparts, self.mode, self.device)
