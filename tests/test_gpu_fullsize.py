"""Full-size GPU checks of BASELINE.json's single-GPU configurations.

Against the CPU oracle (the C restatement of the reference's algorithm, oracle/; it assembles these sizes in
1.5-4 s each, the streams come from its own generators, nothing here reads /root/reference):
  * cfg 2 (P1-FEM 128^3), cfg 3 (fdrand 200^3 build + a values-only re-assembly into the frozen pattern) and
    cfg 4 (block reaction-diffusion 96^3 x 4 + Dirichlet penalty rows + elimination): colptr / rowval / nzval of
    the GPU product path equal the oracle's bit for bit (test/test_assembly.jl:20-32 asks for exact ==).
Size-independent properties on top:
  * two independent implementations of flush! -- grouped chunks + thread-per-column merge (the product path)
    and the (col,row) radix sort + flat segmented fold (STRATEGY_FULLSORT) -- agree bit for bit;
  * closed-form nnz of the stencil, strictly increasing rows per column, structural symmetry;
  * XSB_FAST (plain and with window pre-aggregation) keeps the pattern and stays within 1e-14;
  * the frozen-pattern re-assembly (Newton loop of configs[2]) reproduces a full flush! bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xsb():
    import __graft_entry__ as ge

    ge.build()
    import xsparse_b200

    assert xsparse_b200.capi.device_count() > 0
    return xsparse_b200


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def device_csc(torch, h):
    nnz = h.nnz
    cp = torch.empty(h.n + 1, dtype=torch.int64, device="cuda")
    rv = torch.empty(nnz, dtype=torch.int64, device="cuda")
    nz = torch.empty(nnz, dtype=torch.float64, device="cuda")
    h.fetch_csc(cp, rv, nz)
    h.synchronize()
    return cp, rv, nz


def check_structure(torch, cp, rv, n, nnz):
    assert int(cp[0]) == 1 and int(cp[-1]) == nnz + 1
    cols = torch.repeat_interleave(torch.arange(1, n + 1, device="cuda"), cp[1:] - cp[:-1])
    same_col = cols[1:] == cols[:-1]
    assert bool(torch.all((rv[1:] > rv[:-1]) | ~same_col)), "rows not strictly increasing inside a column"
    # structural symmetry: the multiset of (i,j) equals the multiset of (j,i)
    a = torch.sort(rv * (n + 1) + cols).values
    b = torch.sort(cols * (n + 1) + rv).values
    assert torch.equal(a, b), "pattern not structurally symmetric"
    return cols


def assemble(xsb, n, emit, strategy=None, mode=None, preagg=False):
    h = xsb.Handle(n, n)
    if strategy is not None:
        h.set_strategy(strategy)
    if preagg:
        h.set_preaggregation(True)
    emit(h)
    n_ins = h.pending
    nnz, changed = h.flush(xsb.DETERMINISTIC if mode is None else mode)
    assert changed
    return h, n_ins, nnz


CASES = {
    # name: (unknowns, emitter, inserted, nnz)
    "cfg2_fem128": (128 ** 3, lambda x: (lambda h: h.emit_p1fem(128, 128, 128, flavour=x.RAW)), 245805960, 31065598),
    "cfg3_fd200": (200 ** 3, lambda x: (lambda h: h.emit_fdrand(200, 200, 200, seed=5)), 95760000, 55760000),
    "cfg4_rd96": (4 * 96 ** 3, lambda x: (lambda h: h.emit_blockrd(96, 96, 96, 4, seed=5)), 182255616, 98205696),
}


@pytest.mark.parametrize("case", list(CASES))
def test_two_flush_implementations_agree_bitwise(xsb, torch, case):
    n, mk, n_ins_ref, nnz_ref = CASES[case]
    emit = mk(xsb)
    h, n_ins, nnz = assemble(xsb, n, emit)
    st = h.flush_stats()
    assert (n_ins, nnz) == (n_ins_ref, nnz_ref)
    assert st["column_path"] == 4  # the product path of the bench: grouped chunks + thread-per-column merge
    cp, rv, nz = device_csc(torch, h)
    h.close()
    check_structure(torch, cp, rv, n, nnz)
    g, _, nnz2 = assemble(xsb, n, emit, strategy=xsb.capi.STRATEGY_FULLSORT)
    assert g.flush_stats()["column_path"] == 0 and nnz2 == nnz
    cp2, rv2, nz2 = device_csc(torch, g)
    g.close()
    assert torch.equal(cp, cp2) and torch.equal(rv, rv2)
    assert torch.equal(nz.view(torch.int64), nz2.view(torch.int64)), "nzval differs between the two flush paths"


def test_insertion_paths_agree_with_the_emitter(xsb, torch):
    """The generic insertion paths at a size that exercises their large-batch machinery -- the FEM 80^3 stream
    (59 M rawupdateindex! calls): (a) 16-byte triplets in HOST memory (xsb_insert_triplets: slices over PCIe on a
    second stream, packed in place while the next slice travels), (b) the same triplets resident on the DEVICE
    (grouped through shared-memory keys, re-read in destination order), (c) device (I,J,V) arrays, (d) host (I,J,V)
    arrays -- all give the CSC of the stream generated on the device, bit for bit; a BoundsError in the LAST slice
    rejects the whole batch."""
    mesh = 80
    n = mesh ** 3
    ref = xsb.Handle(n, n)
    ref.set_precount(False)  # the staged records in call order
    ref.emit_p1fem(mesh, mesh, mesh, flavour=xsb.RAW)
    cnt = ref.pending
    assert cnt >= 3 * (1 << 22)
    dI = torch.empty(cnt, dtype=torch.int64, device="cuda")
    dJ = torch.empty_like(dI)
    dV = torch.empty(cnt, dtype=torch.float64, device="cuda")
    got = C.c_int64(0)
    c = xsb.capi
    c.check(c.lib().xsb_debug_fetch_staged(ref._h, 0, dI.data_ptr(), dJ.data_ptr(), dV.data_ptr(), None, cnt,
                                           C.byref(got)), ref._h)
    ref.flush()
    want = [t.clone() for t in device_csc(torch, ref)]
    ref.close()
    dT = torch.empty((cnt, 2), dtype=torch.int64, device="cuda")
    dT[:, 0] = dI | (dJ << 32)
    dT[:, 1] = dV.view(torch.int64)
    hT = dT.cpu()

    def same(h):
        for a, b in zip(device_csc(torch, h), want):
            assert torch.equal(a.view(torch.int64), b.view(torch.int64))

    h = xsb.Handle(n, n)
    h.insert_triplets(hT, xsb.RAW, 0, cnt)  # (a)
    h.flush()
    assert h.flush_stats()["column_path"] == 4
    same(h)
    h.reset()
    h.insert_triplets(dT, xsb.RAW, 0, cnt)  # (b)
    h.flush()
    same(h)
    h.reset()
    h.insert_batch(dI, dJ, dV, xsb.RAW)  # (c)
    h.flush()
    same(h)
    h.reset()
    h.insert_batch(dI.cpu().numpy(), dJ.cpu().numpy(), dV.cpu().numpy(), xsb.RAW)  # (d)
    h.flush()
    same(h)
    h.reset()
    bad = hT.clone()
    bad[cnt - 7, 0] = (n + 1) | (1 << 32)  # row n + 1 in the last slice
    with pytest.raises(IndexError) as e:
        h.insert_triplets(bad, xsb.RAW, 0, cnt)
    assert f"entry {cnt - 7} " in str(e.value)  # position in the caller's batch, not in the slice
    assert h.pending == 0
    h.insert_triplets(hT, xsb.RAW, 0, cnt)
    h.flush()
    same(h)
    h.close()


@pytest.mark.parametrize("preagg", [False, True])
def test_fast_mode_full_size_fem(xsb, torch, preagg):
    n, mk, _, nnz_ref = CASES["cfg2_fem128"]
    emit = mk(xsb)
    h, _, nnz = assemble(xsb, n, emit)
    cp, rv, nz = device_csc(torch, h)
    h.close()
    g, n_ins, nnz2 = assemble(xsb, n, emit, mode=xsb.FAST, preagg=preagg)
    st = g.flush_stats()
    assert nnz2 == nnz == nnz_ref
    if preagg:
        assert 0 < st["preagg_records"] < n_ins // 2
    cp2, rv2, nz2 = device_csc(torch, g)
    g.close()
    assert torch.equal(cp, cp2) and torch.equal(rv, rv2)
    tol = 1e-14 * float(nz.abs().max())
    assert float((nz - nz2).abs().max()) <= tol


def test_newton_loop_fd200_values_only(xsb, torch):
    """configs[2]: fdrand 200^3 built once, then values-only re-assemblies into the frozen pattern;
    every re-assembly must equal a fresh insert + flush! of the same stream bit for bit."""
    nx = 200
    n = nx ** 3
    c = xsb.capi

    def staged(seed):
        g = xsb.Handle(n, n)
        g.set_precount(False)  # staged in stream order: this IS the user's (I, J, V) stream
        g.emit_fdrand(nx, nx, nx, seed=seed)
        cnt = g.pending
        dI = torch.empty(cnt, dtype=torch.int64, device="cuda")
        dJ = torch.empty(cnt, dtype=torch.int64, device="cuda")
        dV = torch.empty(cnt, dtype=torch.float64, device="cuda")
        got = C.c_int64(0)
        c.check(c.lib().xsb_debug_fetch_staged(g._h, 0, dI.data_ptr(), dJ.data_ptr(), dV.data_ptr(), None, cnt,
                                               C.byref(got)), g._h)
        g.synchronize()
        return g, dI, dJ, dV

    g0, dI, dJ, dV = staged(100)
    g0.close()
    h = xsb.Handle(n, n)
    h.insert_batch(dI, dJ, dV, xsb.UPDATE, count=len(dV))
    nnz, _ = h.flush()
    assert nnz == 7 * n - 6 * nx * nx
    h.freeze_pattern(dI, dJ, count=len(dI))
    for seed in (101, 102, 103):
        g, _, _, dVn = staged(seed)
        g.flush()
        _, _, ref = device_csc(torch, g)
        g.close()
        h.zero_values()
        h.reassemble_values(dVn, xsb.DETERMINISTIC, count=len(dVn))
        _, _, nz = device_csc(torch, h)
        assert torch.equal(nz.view(torch.int64), ref.view(torch.int64)), f"re-assembly {seed} not bit-exact"
        h.reassemble_values(dVn, xsb.DETERMINISTIC, count=len(dVn), zero_first=True)  # nonzeros .= 0 fused in
        _, _, nz = device_csc(torch, h)
        assert torch.equal(nz.view(torch.int64), ref.view(torch.int64)), f"zeroed re-assembly {seed} not bit-exact"
        h.zero_values()
        h.reassemble_values(dVn, xsb.FAST, count=len(dVn))
        _, _, nzf = device_csc(torch, h)
        assert float((nzf - ref).abs().max()) <= 1e-14 * float(ref.abs().max())
        del dVn, ref, nz, nzf
    h.close()


# ---------------------------------------------------------------- full size against the oracle
def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _assert_equals_oracle(h, ref, what):
    cp, rv, nz = h.fetch_csc_numpy()
    ocp, orv, onz = ref
    assert np.array_equal(cp, ocp), f"{what}: colptr differs from the oracle"
    assert np.array_equal(rv, orv), f"{what}: rowval differs from the oracle"
    bad = np.nonzero(_bits(nz) != _bits(onz))[0]
    assert bad.size == 0, f"{what}: {bad.size} nzval entries not bit-exact with the oracle, first at {bad[:5]}"


def test_cfg2_fem128_equals_oracle(xsb, oracle):
    """configs[1]: the bench workload.  GPU product path (on-device emitter, grouped chunks) vs the oracle's
    ExtendableSparseMatrix + rawupdateindex! + flush! on the oracle's own copy of the stream."""
    n1 = 128
    n = n1 ** 3
    I, J, V = oracle.fem_stream(n1, n1, n1)
    A = oracle.OracleExt(n, n)
    A.insert_batch(I, J, V, oracle.RAW)
    del I, J, V
    ref = A.csc()
    del A
    h = xsb.Handle(n, n)
    h.emit_p1fem(n1, n1, n1, flavour=xsb.RAW)
    nnz, changed = h.flush()
    assert changed and nnz == len(ref[1]) == 31065598 and h.flush_stats()["column_path"] == 4
    _assert_equals_oracle(h, ref, "cfg2")
    h.close()


def test_cfg3_fd200_build_and_reassembly_equal_oracle(xsb, oracle):
    """configs[2]: fdrand 200^3 via updateindex! + flush!, then one values-only re-assembly (nonzeros .= 0, second
    stream into the frozen pattern: every call a CSC hit, extendable.jl:164-166) -- both against the oracle."""
    nx = 200
    n = nx ** 3
    I, J, V = oracle.fdrand_stream(nx, nx, nx, seed=100)
    A = oracle.OracleExt(n, n)
    A.insert_batch(I, J, V, oracle.UPDATE)
    h = xsb.Handle(n, n)
    h.emit_fdrand(nx, nx, nx, seed=100)
    nnz, _ = h.flush()
    assert nnz == 55760000 and h.flush_stats()["column_path"] == 4
    _assert_equals_oracle(h, A.csc(), "cfg3 build")
    _, _, V2 = oracle.fdrand_stream(nx, nx, nx, seed=101)
    A.zero_values()
    A.insert_batch(I, J, V2, oracle.UPDATE)
    h.freeze_pattern(I, J)
    h.zero_values()
    h.reassemble_values(V2, xsb.DETERMINISTIC)
    _assert_equals_oracle(h, A.csc(), "cfg3 re-assembly")
    h.reassemble_values(V2, xsb.DETERMINISTIC, zero_first=True)
    _assert_equals_oracle(h, A.csc(), "cfg3 re-assembly (zeroed)")
    h.reassemble_values(V, xsb.DETERMINISTIC)  # on top of the resident values: ((old + v1) + v2) ...
    A.insert_batch(I, J, V, oracle.UPDATE)
    _assert_equals_oracle(h, A.csc(), "cfg3 accumulate onto resident values")
    h.close()


def test_cfg4_rd96_with_dirichlet_equals_oracle(xsb, oracle):
    """configs[3]: 96^3 x 4-species block system, Dirichlet penalty rows on the two x-faces (A[d,d] = 1e30, the
    assign flavour, test/test_dirichlet.jl:9-11) and eliminate_dirichlet! (sparsematrixcsc.jl:124-148)."""
    nx, ns = 96, 4
    n = ns * nx ** 3
    I, J, V = oracle.blockrd_stream(nx, nx, nx, ns, seed=7)
    A = oracle.OracleExt(n, n)
    A.insert_batch(I, J, V, oracle.UPDATE)
    del I, J, V
    node = np.arange(nx ** 3)
    face = (node % nx == 0) | (node % nx == nx - 1)
    d = (ns * node[face][:, None] + np.arange(ns)[None, :]).reshape(-1) + 1
    pen = np.full(len(d), 1.0e30)
    A.insert_batch(d, d, pen, oracle.ASSIGN)
    A.flush()
    h = xsb.Handle(n, n)
    h.emit_blockrd(nx, nx, nx, ns, seed=7, flavour=xsb.UPDATE)
    h.insert_batch(d, d, pen, xsb.ASSIGN)
    nnz, _ = h.flush()
    assert nnz == 98205696
    _assert_equals_oracle(h, A.csc(), "cfg4 assembly")
    mk = h.mark_dirichlet(1.0e20)
    assert np.array_equal(mk, A.mark_dirichlet(1.0e20)) and int(mk.sum()) == len(d)
    h.eliminate_dirichlet(mk)
    A.eliminate_dirichlet(mk)
    _assert_equals_oracle(h, A.csc(), "cfg4 eliminated")
    h.close()
