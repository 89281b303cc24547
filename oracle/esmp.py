"""CPU restatement of the reference's experimental ExtendableSparseMatrixParallel (test infrastructure only).

The module is disabled in ExtendableSparse.jl v1.5.1 (`src/ExtendableSparse.jl:50` is commented out), but its
assembly path is part of SURVEY.md 8(a) row a16.  Restated here, in plain Python loops (small cases only):

  * per-thread buffers with LOCAL column numbering: `sortednodesperthread[tid, j]` = local column of global column j
    in thread tid's buffer (0: not owned), `globalindices[tid][local]` = global column
    (`src/experimental/ExtendableSparseMatrixParallel/ExtendableSparseParallel.jl:5-92`);
  * `addtoentry!(A, i, j, tid, v)` (`ExtendableSparseParallel.jl:125-135`): a CSC hit is added in place
    (`updatentryCSC2!`, `:388-400`), a miss becomes `lnkmatrices[tid][i, local(j)] += v`, i.e. getindex + setindex! on
    the SuperSparseMatrixLNK (`supersparse.jl:125-156`): a new entry is only created when the stored value is non-zero;
  * `flush!` (`struct_flush.jl:1-32`), sparse path = `plus_remap` (`supersparse.jl:408-514`): the touched GLOBAL columns
    in ascending order; per column the lists of all owning threads are gathered in thread order, sorted by row
    (`get_column_keepzeros!`, `:270-292`), equal rows summed (`remove_doubles!`, `:230-244`) and merged into the old CSC
    column (`merge_into!`, `:294-389`); untouched column ranges are copied.

NOT copied: `flush!` computes `A.cscmatrix + plus_remap(lnks, A.cscmatrix, ...)` (`struct_flush.jl:10`),
which counts the values of a non-empty old CSC twice; the intended result `csc + sum(lnks)` is restated.  The
reference sorts with an unstable QuickSort, so the order in which the threads' partial sums of one entry are added is
not defined there: values are compared with `isapprox` (as `test/ExperimentalParallel.jl:287` does), the pattern exactly.
"""
from __future__ import annotations

import numpy as np


class ESMP:
    """ExtendableSparseMatrixParallel{Float64,Int64} on an n x n matrix with `nt` thread buffers; column j is owned
    by the threads listed in `owners[j]` (1-based thread ids; a separator column may have several owners)."""

    def __init__(self, n: int, nt: int, owners):
        self.n, self.nt = int(n), int(nt)
        # sortednodesperthread[tid][j] (0 = not owned) and globalindices[tid] (ExtendableSparseParallel.jl:40-60)
        self.sortednodesperthread = np.zeros((nt + 1, n + 1), np.int64)
        self.globalindices = [[] for _ in range(nt + 1)]
        for j in range(1, n + 1):
            for tid in owners[j - 1]:
                self.globalindices[tid].append(j)
                self.sortednodesperthread[tid, j] = len(self.globalindices[tid])
        self.colptr = np.ones(n + 1, np.int64)
        self.rowval = np.empty(0, np.int64)
        self.nzval = np.empty(0, np.float64)
        self._new_buffers()

    def _new_buffers(self):
        # lnkmatrices[tid]: {local column: [(row, value), ...] in order of first insertion}
        self.lnk = [dict() for _ in range(self.nt + 1)]

    def _csc_find(self, i, j):
        a, b = self.colptr[j - 1] - 1, self.colptr[j] - 1
        k = a + np.searchsorted(self.rowval[a:b], i)
        return int(k) if k < b and self.rowval[k] == i else -1

    def addtoentry(self, i, j, tid, v):
        """addtoentry!(A, i, j, tid, v): ExtendableSparseParallel.jl:125-135."""
        k = self._csc_find(i, j)
        if k >= 0:  # updatentryCSC2!: in place
            self.nzval[k] += v
            return
        loc = int(self.sortednodesperthread[tid, j])
        assert loc > 0, "column not owned by this thread"
        col = self.lnk[tid].setdefault(loc, [])
        for e in col:  # getindex + setindex!: supersparse.jl:125-156
            if e[0] == i:
                e[1] = e[1] + v
                return
        if v != 0.0:  # a new entry only for a non-zero value
            col.append([i, 0.0 + v])

    def nnz_lnk(self):
        return sum(len(c) for b in self.lnk for c in b.values())

    def flush(self):
        """flush!(A; do_dense=false, keep_zeros=true): struct_flush.jl:1-32 -> plus_remap, supersparse.jl:408-514."""
        if self.nnz_lnk() == 0:
            return
        touched = {}  # global column -> [(tid, local column)] in thread order
        for tid in range(1, self.nt + 1):
            for loc in self.lnk[tid]:
                if self.lnk[tid][loc]:
                    touched.setdefault(self.globalindices[tid][loc - 1], []).append((tid, loc))
        colptr = [1]
        rowval, nzval = [], []
        for j in range(1, self.n + 1):
            a, b = self.colptr[j - 1] - 1, self.colptr[j] - 1
            old = list(zip(self.rowval[a:b].tolist(), self.nzval[a:b].tolist()))
            if j in touched:
                col = [tuple(e) for (tid, loc) in touched[j] for e in self.lnk[tid][loc]]  # get_column_keepzeros!
                col.sort(key=lambda e: e[0])  # (the reference's QuickSort is not stable)
                merged = []
                for r, v in col:  # remove_doubles!
                    if merged and merged[-1][0] == r:
                        merged[-1][1] += v
                    else:
                        merged.append([r, v])
                out, p, q = [], 0, 0  # merge_into!
                while p < len(old) or q < len(merged):
                    if q == len(merged) or (p < len(old) and old[p][0] < merged[q][0]):
                        out.append(old[p])
                        p += 1
                    elif p == len(old) or merged[q][0] < old[p][0]:
                        out.append(tuple(merged[q]))
                        q += 1
                    else:
                        out.append((old[p][0], old[p][1] + merged[q][1]))
                        p += 1
                        q += 1
                old = out
            rowval += [r for r, _ in old]
            nzval += [v for _, v in old]
            colptr.append(len(rowval) + 1)
        self.colptr = np.asarray(colptr, np.int64)
        self.rowval = np.asarray(rowval, np.int64)
        self.nzval = np.asarray(nzval, np.float64)
        self._new_buffers()

    def csc(self):
        return self.colptr.copy(), self.rowval.copy(), self.nzval.copy()
