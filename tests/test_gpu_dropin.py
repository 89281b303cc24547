"""The drop-in boundary, executed: the Python mirror of the Julia glue (extendablesparse.jl_b200/dropin.py)
drives libxsparse_b200 with exactly the call sequence julia/ExtendableSparseB200.jl makes
(xsb_set_csc -> xsb_insert_triplets per partition and flavour run -> xsb_flush -> xsb_fetch_csc, one handle),
behind a restatement of the reference's own wrapper (genericmtextendablesparsematrixcsc.jl:1-114), and is
compared bit for bit with the oracle's MTExtendableSparseMatrixCSC.  Also: the header's threading contract
(concurrent insertions with distinct tid) and the MT wrapper's error contract."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xsb():
    import __graft_entry__ as ge

    ge.build()
    import xsparse_b200

    assert xsparse_b200.capi.device_count() > 0, "no CUDA device: GPU tests need the real extension on a GPU"
    return xsparse_b200


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def assert_csc_equal(got, ref):
    assert np.array_equal(got[0], ref[0]), "colptr differs"
    assert np.array_equal(got[1], ref[1]), "rowval differs"
    bad = np.nonzero(bits(got[2]) != bits(ref[2]))[0]
    assert bad.size == 0, f"{bad.size} nzval entries not bit-exact, first at {bad[:5]}"


@pytest.mark.parametrize("nparts", [1, 3])
def test_dropin_wrapper_matches_oracle_mt(xsb, oracle, nparts):
    """test/test_assembly.jl:6-35 through the MT wrapper: random insertions with duplicates, several splices;
    hits are folded on the host (the wrapper's CSC branch), misses travel through Base.sum(exts, csc)."""
    from xsparse_b200 import dropin

    rng = np.random.default_rng(1234 + nparts)
    m, n = 700, 500
    A = dropin.GenericMTExtendableSparseMatrixCSC(m, n, nparts)
    M = oracle.OracleMT(m, n, nparts)
    for splice in range(4):
        for t in range(nparts):
            cnt = int(rng.integers(2000, 6000))
            I = rng.integers(1, m + 1, cnt)
            J = rng.integers(1, n + 1, cnt)
            V = rng.standard_normal(cnt)
            V[rng.random(cnt) < 0.05] = 0.0
            fl = (xsb.RAW, xsb.UPDATE)[(splice + t) % 2]
            A.update_batch(fl, I, J, V, tid=t + 1)
            M.insert_batch(I, J, V, t + 1, {xsb.RAW: oracle.RAW, xsb.UPDATE: oracle.UPDATE}[fl])
        # a few per-entry calls as well (the path a Julia loop takes)
        for _ in range(50):
            i, j, v = int(rng.integers(1, m + 1)), int(rng.integers(1, n + 1)), float(rng.standard_normal())
            t = int(rng.integers(1, nparts + 1))
            A.rawupdateindex("+", v, i, j, t)
            M.update(v, i, j, t, oracle.RAW)
        A.flush()
        M.flush()
        assert_csc_equal(A.sparse(), M.csc())
    assert A.pattern_changes >= 1
    # setindex! of an existing entry works in place, of a new entry is the reference's error
    cp, rv, nz = A.sparse()
    j0 = int(np.nonzero(np.diff(cp))[0][0]) + 1
    i0 = int(rv[cp[j0 - 1] - 1])
    A[i0, j0] = 3.25
    assert A[i0, j0] == 3.25
    free = next((i, j0) for i in range(1, m + 1) if i not in set(rv[cp[j0 - 1] - 1:cp[j0] - 1].tolist()))
    with pytest.raises(xsb.capi.XsbIllegalError):
        A[free[0], free[1]] = 1.0
    dropin.release_handles()


def test_dropin_sum_sequence_fem(xsb, oracle):
    """Base.sum on the FEM stream: build from an empty CSC, then a second assembly on top of the first
    (every call a hit -> nothing staged -> flush! is a no-op) and a splice of new entries."""
    from xsparse_b200 import dropin

    n1 = 14
    I, J, V = oracle.fem_stream(n1, n1, n1)
    n = n1 ** 3
    A = dropin.GenericMTExtendableSparseMatrixCSC(n, n, 1)
    M = oracle.OracleMT(n, n, 1)
    A.update_batch(xsb.RAW, I, J, V)
    M.insert_batch(I, J, V, 1, oracle.RAW)
    A.flush()
    M.flush()
    assert_csc_equal(A.sparse(), M.csc())
    A.update_batch(xsb.RAW, I, J, V)  # all hits
    M.insert_batch(I, J, V, 1, oracle.RAW)
    assert A.nnznew == 0
    A.flush()
    M.flush()
    assert_csc_equal(A.sparse(), M.csc())
    rng = np.random.default_rng(3)
    I2, J2 = rng.integers(1, n + 1, 3000), rng.integers(1, n + 1, 3000)
    V2 = rng.standard_normal(3000)
    A.update_batch(xsb.UPDATE, I2, J2, V2)
    M.insert_batch(I2, J2, V2, 1, oracle.UPDATE)
    A.flush()
    M.flush()
    assert_csc_equal(A.sparse(), M.csc())
    dropin.release_handles()


def test_concurrent_inserts_with_distinct_tid(xsb, oracle):
    """include/xsparse_b200.h: xsb_insert_* may be called concurrently with DISTINCT tid (the reference's
    threading contract, test/femtools.jl:88-105).  8 host threads insert into one handle at once; the result
    equals the oracle's partition-ordered flush, and a BoundsError in one thread's batch is reported to it."""
    nparts, m, n = 8, 900, 900
    rng = np.random.default_rng(99)
    parts = []
    for t in range(nparts):
        cnt = 20000 + 1000 * t
        parts.append((rng.integers(1, m + 1, cnt), rng.integers(1, n + 1, cnt), rng.standard_normal(cnt)))
    M = oracle.OracleMT(m, n, nparts)
    for t, (I, J, V) in enumerate(parts):
        M.insert_batch(I, J, V, t + 1, oracle.RAW)
    M.flush()
    for attempt in range(3):
        h = xsb.Handle(m, n, n_tid=nparts)
        errors = [None] * nparts
        start = threading.Barrier(nparts)

        def work(t):
            I, J, V = parts[t]
            try:
                start.wait()
                for a in range(0, len(V), 3000):  # many small calls per thread: plenty of interleaving
                    h.insert_batch(I[a:a + 3000], J[a:a + 3000], V[a:a + 3000], xsb.RAW, tid=t)
                if t == 5:  # a rejected batch must not disturb the other threads' batches
                    try:
                        h.insert_batch(np.array([m + 1]), np.array([1]), np.array([1.0]), xsb.RAW, tid=t)
                    except IndexError:
                        errors[t] = "bounds"
            except Exception as e:  # noqa: BLE001
                errors[t] = e

        threads = [threading.Thread(target=work, args=(t,)) for t in range(nparts)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        assert errors[5] == "bounds" and all(e is None for k, e in enumerate(errors) if k != 5), errors
        h.flush()
        assert_csc_equal(h.fetch_csc_numpy(), M.csc())
        h.close()


def test_mt_setindex_of_new_entry_is_illegal(xsb):
    """genericmtextendablesparsematrixcsc.jl:63-68: A[i,j] = v on the multi-partition wrapper is an error unless
    the entry is in the CSC already.  XSB_EILLEGAL rejects the batch; the handle stays usable."""
    h = xsb.Handle(20, 20, n_tid=2)
    h.insert_batch(np.array([1, 2, 3]), np.array([1, 2, 3]), np.array([1.0, 2.0, 3.0]), xsb.RAW, tid=0)
    h.flush()
    with pytest.raises(xsb.capi.XsbIllegalError) as e:
        h.insert_batch(np.array([2, 4]), np.array([2, 4]), np.array([9.0, 9.0]), xsb.ASSIGN, tid=1)
    assert "rawupdateindex" in str(e.value) and h.pending == 0
    h.insert_batch(np.array([2]), np.array([2]), np.array([9.0]), xsb.ASSIGN, tid=1)  # existing entry: legal
    T = np.zeros(1, xsb.capi.TRIPLET_DTYPE)
    T[0] = (3, 3, 7.0)
    h.insert_triplets(T, xsb.ASSIGN, 0)
    T[0] = (5, 6, 7.0)
    with pytest.raises(xsb.capi.XsbIllegalError):
        h.insert_triplets(T, xsb.ASSIGN, 0)
    assert h.flush() == (3, False)
    assert h.fetch_csc_numpy()[2].tolist() == [1.0, 9.0, 7.0]
    # a single-partition handle takes A[i,j] = v for new entries (extendable.jl:205-218)
    g = xsb.Handle(20, 20)
    g.insert_batch(np.array([4]), np.array([4]), np.array([9.0]), xsb.ASSIGN)
    assert g.flush() == (1, True)
    h.close()
    g.close()


def test_esmp_assembly_through_partition_buffers(xsb):
    """the experimental ExtendableSparseMatrixParallel (addtoentry! into per-thread buffers with
    local column numbering, flush! = plus_remap) restated in oracle/esmp.py, against the library's multi-partition
    handle fed the same calls (GLOBAL columns, tid = thread): the pattern must be identical, the values `≈`
    (test/ExperimentalParallel.jl:287; the reference's own summation order across threads is not defined)."""
    from oracle.esmp import ESMP
    from xsparse_b200 import dropin

    rng = np.random.default_rng(21)
    n, nt = 80, 4
    owners = []
    for j in range(n):
        t = j * nt // n + 1
        owners.append([t] if j % 5 else sorted({t, t % nt + 1}))
    A = ESMP(n, nt, owners)
    G = dropin.GenericMTExtendableSparseMatrixCSC(n, n, nt)
    for splice in range(3):
        for _ in range(2500):
            j = int(rng.integers(1, n + 1))
            tid = int(rng.choice(owners[j - 1]))
            i = int(rng.integers(1, n + 1))
            v = float(rng.standard_normal()) if rng.random() > 0.05 else 0.0
            A.addtoentry(i, j, tid, v)
            G.updateindex("+", v, i, j, tid)  # A[i,j] += v creates an entry only for a non-zero value: updateindex!
        A.flush()
        G.flush()
        cp, rv, nz = G.sparse()
        ocp, orv, onz = A.csc()
        assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
        assert np.allclose(nz, onz, rtol=1e-13, atol=1e-14)
    dropin.release_handles()
