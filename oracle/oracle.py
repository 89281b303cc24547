"""ctypes front end of the CPU ORACLE (oracle/xsb_oracle.c, oracle/xsb_streams.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.

All indices are 1-based, as in the Julia reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libxsb_oracle.so")

UPDATE, RAW, ASSIGN = 0, 1, 2

_i64 = C.c_int64
_f64 = C.c_double
_p = C.c_void_p


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("xsb_oracle.c", "xsb_streams.c", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s)
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)

        def sig(name, res, *args):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = list(args)

        sig("ora_lnk_create", _p, _i64, _i64)
        sig("ora_lnk_destroy", None, _p)
        sig("ora_lnk_nnz", _i64, _p)
        sig("ora_lnk_setindex", C.c_int, _p, _f64, _i64, _i64)
        sig("ora_lnk_updateindex", C.c_int, _p, _f64, _i64, _i64)
        sig("ora_lnk_rawupdateindex", C.c_int, _p, _f64, _i64, _i64)
        sig("ora_lnk_getindex", C.c_int, _p, _i64, _i64, C.POINTER(_f64))
        sig("ora_lnk_plus_csc", _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p)
        sig("ora_ext_create", _p, _i64, _i64)
        sig("ora_ext_destroy", None, _p)
        sig("ora_ext_reset", None, _p)
        sig("ora_ext_set_csc", None, _p, _p, _p, _p)
        sig("ora_ext_updateindex", C.c_int, _p, _f64, _i64, _i64)
        sig("ora_ext_rawupdateindex", C.c_int, _p, _f64, _i64, _i64)
        sig("ora_ext_setindex", C.c_int, _p, _f64, _i64, _i64)
        sig("ora_ext_getindex", C.c_int, _p, _i64, _i64, C.POINTER(_f64))
        sig("ora_ext_flush", None, _p)
        sig("ora_ext_nflush", _i64, _p)
        sig("ora_ext_nnz", _i64, _p)
        sig("ora_ext_nnz_csc", _i64, _p)
        sig("ora_ext_nnz_lnk", _i64, _p)
        sig("ora_ext_get_csc", None, _p, _p, _p, _p)
        sig("ora_ext_zero_values", None, _p)
        sig("ora_ext_insert_batch", _i64, _p, _p, _p, _p, _i64, C.c_int)
        sig("ora_ext_mark_dirichlet", None, _p, _f64, _p)
        sig("ora_ext_eliminate_dirichlet", None, _p, _p)
        sig("ora_ext_pointblock", _i64, _p, _i64, _p, _p, _p)
        sig("ora_mt_create", _p, _i64, _i64, _i64)
        sig("ora_mt_destroy", None, _p)
        sig("ora_mt_update", C.c_int, _p, _f64, _i64, _i64, _i64, C.c_int)
        sig("ora_mt_insert_batch", _i64, _p, _p, _p, _p, _i64, _i64, C.c_int)
        sig("ora_mt_insert_partitioned", _i64, _p, _p, _p, _p, _p, _i64, C.c_int)
        sig("ora_mt_flush", None, _p)
        sig("ora_mt_nnz", _i64, _p)
        sig("ora_mt_nnznew", _i64, _p)
        sig("ora_mt_get_csc", None, _p, _p, _p, _p)
        sig("ora_mt_zero_values", None, _p)
        sig("ora_philox4x32_10", None, C.c_uint64, C.c_uint64, _p)
        sig("ora_uniform", _f64, C.c_uint64, C.c_uint64)
        sig("ora_fdrand_count", _i64, _i64, _i64, _i64)
        sig("ora_fdrand_stream", None, _i64, _i64, _i64, C.c_uint64, C.c_int, _p, _p, _p)
        sig("ora_fem_count", _i64, _i64, _i64, _i64)
        sig("ora_fem_stream", None, _i64, _i64, _i64, _p, _p, _p)
        sig("ora_blockrd_count", _i64, _i64, _i64, _i64, _i64)
        sig("ora_blockrd_stream", None, _i64, _i64, _i64, _i64, C.c_uint64, _p, _p, _p)
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_p)


def _i64a(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.int64)


def _f64a(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float64)


class OracleBoundsError(IndexError):
    pass


class OracleLNK:
    """SparseMatrixLNK (src/matrix/sparsematrixlnk.jl)."""

    def __init__(self, m: int, n: int):
        self.m, self.n = int(m), int(n)
        self._h = lib().ora_lnk_create(self.m, self.n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_lnk_destroy(self._h)
            self._h = None

    @classmethod
    def from_csc(cls, m, n, colptr, rowval, nzval):
        """SparseMatrixLNK(csc): sparsematrixlnk.jl:109-118 (setindex! per entry)."""
        L = cls(m, n)
        for j in range(1, n + 1):
            for k in range(int(colptr[j - 1]), int(colptr[j])):
                L[int(rowval[k - 1]), j] = float(nzval[k - 1])
        return L

    @property
    def nnz(self):
        return lib().ora_lnk_nnz(self._h)

    def _chk(self, rc):
        if rc:
            raise OracleBoundsError("BoundsError")

    def __setitem__(self, ij, v):
        self._chk(lib().ora_lnk_setindex(self._h, float(v), int(ij[0]), int(ij[1])))

    def __getitem__(self, ij):
        out = _f64()
        self._chk(lib().ora_lnk_getindex(self._h, int(ij[0]), int(ij[1]), C.byref(out)))
        return out.value

    def updateindex(self, v, i, j):
        self._chk(lib().ora_lnk_updateindex(self._h, float(v), int(i), int(j)))

    def rawupdateindex(self, v, i, j):
        self._chk(lib().ora_lnk_rawupdateindex(self._h, float(v), int(i), int(j)))

    def plus_csc(self, m, n, colptr, rowval, nzval):
        """lnk + csc -> (colptr, rowval, nzval), sparsematrixlnk.jl:294-383."""
        colptr, rowval, nzval = _i64a(colptr), _i64a(rowval), _f64a(nzval)
        cap = int(colptr[n] - 1) + self.nnz
        cp = np.empty(n + 1, np.int64)
        rv = np.empty(max(cap, 1), np.int64)
        nz = np.empty(max(cap, 1), np.float64)
        r = lib().ora_lnk_plus_csc(self._h, m, n, _ptr(colptr), _ptr(rowval), _ptr(nzval),
                                   _ptr(cp), _ptr(rv), _ptr(nz))
        if r < 0:
            raise AssertionError("size mismatch")
        return cp, rv[:r].copy(), nz[:r].copy()


class OracleExt:
    """ExtendableSparseMatrixCSC{Float64,Int64} (src/matrix/extendable.jl)."""

    def __init__(self, m: int, n: int):
        self.m, self.n = int(m), int(n)
        self._h = lib().ora_ext_create(self.m, self.n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_ext_destroy(self._h)
            self._h = None

    def _chk(self, rc):
        if rc:
            raise OracleBoundsError("BoundsError")

    def updateindex(self, v, i, j):
        self._chk(lib().ora_ext_updateindex(self._h, float(v), int(i), int(j)))

    def rawupdateindex(self, v, i, j):
        self._chk(lib().ora_ext_rawupdateindex(self._h, float(v), int(i), int(j)))

    def __setitem__(self, ij, v):
        self._chk(lib().ora_ext_setindex(self._h, float(v), int(ij[0]), int(ij[1])))

    def __getitem__(self, ij):
        out = _f64()
        self._chk(lib().ora_ext_getindex(self._h, int(ij[0]), int(ij[1]), C.byref(out)))
        return out.value

    def insert_batch(self, I, J, V, flavour=UPDATE):
        I, J, V = _i64a(I), _i64a(J), _f64a(V)
        r = lib().ora_ext_insert_batch(self._h, _ptr(I), _ptr(J), _ptr(V), len(V), int(flavour))
        if r < 0:
            raise OracleBoundsError(f"BoundsError at entry {-r - 1}")

    def flush(self):
        lib().ora_ext_flush(self._h)
        return self

    def reset(self):
        lib().ora_ext_reset(self._h)

    def zero_values(self):
        lib().ora_ext_zero_values(self._h)

    def set_csc(self, colptr, rowval, nzval):
        colptr, rowval, nzval = _i64a(colptr), _i64a(rowval), _f64a(nzval)
        lib().ora_ext_set_csc(self._h, _ptr(colptr), _ptr(rowval), _ptr(nzval))

    @property
    def nflush(self):
        return lib().ora_ext_nflush(self._h)

    @property
    def nnz(self):
        return lib().ora_ext_nnz(self._h)

    @property
    def nnz_lnk(self):
        return lib().ora_ext_nnz_lnk(self._h)

    def csc(self, flush=True):
        """(colptr, rowval, nzval), 1-based, after flush! (sparse(A))."""
        if flush:
            self.flush()
        nnz = lib().ora_ext_nnz_csc(self._h)
        cp = np.empty(self.n + 1, np.int64)
        rv = np.empty(nnz, np.int64)
        nz = np.empty(nnz, np.float64)
        lib().ora_ext_get_csc(self._h, _ptr(cp), _ptr(rv) if nnz else None, _ptr(nz) if nnz else None)
        return cp, rv, nz

    def mark_dirichlet(self, penalty=1.0e20):
        mk = np.zeros(self.n, np.uint8)
        lib().ora_ext_mark_dirichlet(self._h, float(penalty), _ptr(mk))
        return mk

    def eliminate_dirichlet(self, marker):
        mk = np.ascontiguousarray(marker, dtype=np.uint8)
        lib().ora_ext_eliminate_dirichlet(self._h, _ptr(mk))


def _ext_pointblock(self, bs):
    """pointblock(A, bs) (extendable.jl:292-318): (colptr, rowval, blocks[nnzb, bs*bs] column-major), 1-based."""
    bs = int(bs)
    cap = max(1, self.nnz)
    cp = np.empty(self.n // bs + 1, np.int64)
    rv = np.empty(cap, np.int64)
    bl = np.zeros(cap * bs * bs, np.float64)
    r = lib().ora_ext_pointblock(self._h, bs, _ptr(cp), _ptr(rv), _ptr(bl))
    if r < 0:
        raise OracleBoundsError(f"entry {-r - 1} falls outside the block matrix")
    return cp, rv[:r].copy(), bl[: r * bs * bs].reshape(r, bs * bs).copy()


OracleExt.pointblock = _ext_pointblock


class OracleMT:
    """GenericMTExtendableSparseMatrixCSC{SparseMatrixDILNKC} (MTExtendableSparseMatrixCSC)."""

    def __init__(self, m: int, n: int, nparts: int = 1):
        self.m, self.n, self.np = int(m), int(n), int(nparts)
        self._h = lib().ora_mt_create(self.m, self.n, self.np)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_mt_destroy(self._h)
            self._h = None

    def update(self, v, i, j, tid=1, flavour=RAW):
        rc = lib().ora_mt_update(self._h, float(v), int(i), int(j), int(tid), int(flavour))
        if rc == 3:
            raise RuntimeError("use rawupdateindex! for new entries")
        if rc:
            raise OracleBoundsError("BoundsError")

    def insert_batch(self, I, J, V, tid=1, flavour=RAW):
        I, J, V = _i64a(I), _i64a(J), _f64a(V)
        r = lib().ora_mt_insert_batch(self._h, _ptr(I), _ptr(J), _ptr(V), len(V), int(tid), int(flavour))
        if r < 0:
            raise OracleBoundsError(f"error at entry {-r - 1}")

    def insert_partitioned(self, I, J, V, part_begin, nthreads=1, flavour=RAW):
        """testassemble_parallel! (test/femtools.jl:75-110): partition p (tid p+1) inserts the slice
        [part_begin[p], part_begin[p+1]) of the stream; two colours (parity of p) run one after the
        other, the partitions of a colour on `nthreads` POSIX threads."""
        I, J, V, pb = _i64a(I), _i64a(J), _f64a(V), _i64a(part_begin)
        assert len(pb) == self.np + 1 and pb[0] >= 0 and pb[-1] <= len(V) and np.all(np.diff(pb) >= 0)
        r = lib().ora_mt_insert_partitioned(self._h, _ptr(I), _ptr(J), _ptr(V), _ptr(pb), int(nthreads), int(flavour))
        if r < 0:
            raise OracleBoundsError(f"error at entry {-r - 1}")

    def flush(self):
        lib().ora_mt_flush(self._h)
        return self

    def zero_values(self):
        lib().ora_mt_zero_values(self._h)

    @property
    def nnznew(self):
        return lib().ora_mt_nnznew(self._h)

    def csc(self):
        nnz = lib().ora_mt_nnz(self._h)
        cp = np.empty(self.n + 1, np.int64)
        rv = np.empty(nnz, np.int64)
        nz = np.empty(nnz, np.float64)
        lib().ora_mt_get_csc(self._h, _ptr(cp), _ptr(rv) if nnz else None, _ptr(nz) if nnz else None)
        return cp, rv, nz


# ---------------------------------------------------------------- streams
def philox(seed: int, counter: int) -> np.ndarray:
    out = np.zeros(4, np.uint32)
    lib().ora_philox4x32_10(seed, counter, _ptr(out))
    return out


def uniform(seed: int, counter: int) -> float:
    return lib().ora_uniform(seed, counter)


def fdrand_count(nx, ny=1, nz=1) -> int:
    return lib().ora_fdrand_count(nx, ny, nz)


def fdrand_stream(nx, ny=1, nz=1, seed=20240717, ones=False):
    """(I, J, V) of fdrand!(A,nx,ny,nz) in call order (sprand.jl:58-126)."""
    cnt = fdrand_count(nx, ny, nz)
    I = np.empty(cnt, np.int64)
    J = np.empty(cnt, np.int64)
    V = np.empty(cnt, np.float64)
    lib().ora_fdrand_stream(nx, ny, nz, seed, int(bool(ones)), _ptr(I), _ptr(J), _ptr(V))
    return I, J, V


def fem_count(nxn, nyn, nzn) -> int:
    return lib().ora_fem_count(nxn, nyn, nzn)


def fem_stream(nxn, nyn, nzn):
    """(I, J, V) of testassemble! on the Kuhn tensor mesh (femtools.jl:45-72)."""
    cnt = fem_count(nxn, nyn, nzn)
    I = np.empty(cnt, np.int64)
    J = np.empty(cnt, np.int64)
    V = np.empty(cnt, np.float64)
    lib().ora_fem_stream(nxn, nyn, nzn, _ptr(I), _ptr(J), _ptr(V))
    return I, J, V


def blockrd_count(nx, ny, nz, ns=4) -> int:
    return lib().ora_blockrd_count(nx, ny, nz, ns)


def blockrd_stream(nx, ny, nz, ns=4, seed=20240717):
    cnt = blockrd_count(nx, ny, nz, ns)
    I = np.empty(cnt, np.int64)
    J = np.empty(cnt, np.int64)
    V = np.empty(cnt, np.float64)
    lib().ora_blockrd_stream(nx, ny, nz, ns, seed, _ptr(I), _ptr(J), _ptr(V))
    return I, J, V
