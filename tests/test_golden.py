"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the CUDA path must match them bit for bit (GPU)."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def replay(d, insert, flush_and_fetch):
    for k in range(int(d["nbatches"])):
        insert(d[f"I{k}"], d[f"J{k}"], d[f"V{k}"], int(d[f"fl{k}"]))
        if bool(d[f"flush{k}"]):
            cp, rv, nz = flush_and_fetch()
            assert np.array_equal(cp, d[f"colptr{k}"]), f"colptr differs after batch {k}"
            assert np.array_equal(rv, d[f"rowval{k}"]), f"rowval differs after batch {k}"
            assert np.array_equal(nz.view(np.uint64), d[f"nzval{k}"].view(np.uint64)), f"nzval differs after batch {k}"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(oracle, path):
    d = np.load(path)
    A = oracle.OracleExt(int(d["m"]), int(d["n"]))
    replay(d, A.insert_batch, A.csc)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_matches_golden(path):
    import __graft_entry__ as ge

    ge.build()
    import xsparse_b200 as xsb

    d = np.load(path)
    h = xsb.Handle(int(d["m"]), int(d["n"]))

    def flush_and_fetch():
        h.flush(xsb.DETERMINISTIC)
        return h.fetch_csc_numpy()

    replay(d, h.insert_batch, flush_and_fetch)


def test_golden_present():
    assert len(GOLDEN) >= 4


# ---------------------------------------------------------------- digests of product-path-sized cases
def _hashes():
    import json

    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hashes.json")))


def _digest(*arrays):
    import hashlib

    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", sorted(_hashes()))
def test_oracle_reproduces_hashed_golden(oracle, name):
    """tests/golden/hashes.json: streams big enough for the flush's product path, recorded as SHA-256."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rec = _hashes()[name]
    got = mg.hashed_case(rec)
    for key in ("n", "records", "nnz", "csc", "pointblock4", "nnz_blocks"):
        assert got[key] == rec[key], key


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_hashes()))
def test_cuda_matches_hashed_golden(name):
    """The library's on-device emitters + flush! (grouping by column, thread fold) + pointblock must reach the
    recorded bytes: colptr, rowval, nzval and the 4x4 point blocks."""
    import __graft_entry__ as ge

    ge.build()
    import xsparse_b200 as xsb

    rec = _hashes()[name]
    h = xsb.Handle(rec["n"], rec["n"])
    fl = xsb.RAW if rec["flavour"] == "raw" else xsb.UPDATE
    if rec["kind"] == "fem":
        h.emit_p1fem(*rec["dims"], flavour=fl)
    elif rec["kind"] == "fd":
        h.emit_fdrand(*rec["dims"], seed=rec["seed"], flavour=fl)
    else:
        h.emit_blockrd(*rec["dims"], seed=rec["seed"], flavour=fl)
    assert h.pending == rec["records"]
    nnz, _ = h.flush(xsb.DETERMINISTIC)
    assert nnz == rec["nnz"] and h.flush_stats()["column_path"] == 4
    assert _digest(*h.fetch_csc_numpy()) == rec["csc"]
    hb = h.pointblock(4)
    assert hb.nnz == rec["nnz_blocks"]
    cp, rv, _ = hb.fetch_csc_numpy()
    assert _digest(cp, rv, hb.fetch_blocks_numpy().reshape(hb.nnz, 16)) == rec["pointblock4"]
    hb.close()
    h.close()
