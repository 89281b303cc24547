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
