"""Generates the committed golden fixtures from the CPU oracle (pinned by tests/test_oracle_kat.py).

    python tests/golden/make_golden.py

The Julia reference cannot run in the build image; these vectors are oracle outputs, recorded so
that (a) the oracle cannot drift silently and (b) the CUDA path is checked against fixed bytes."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as ora  # noqa: E402


def case_stream(name, m, n, batches):
    A = ora.OracleExt(m, n)
    out = {"m": m, "n": n, "nbatches": len(batches)}
    for k, (I, J, V, fl, flush) in enumerate(batches):
        A.insert_batch(I, J, V, fl)
        out[f"I{k}"], out[f"J{k}"], out[f"V{k}"], out[f"fl{k}"], out[f"flush{k}"] = I, J, V, fl, flush
        if flush:
            cp, rv, nz = A.csc()
            out[f"colptr{k}"], out[f"rowval{k}"], out[f"nzval{k}"] = cp, rv, nz
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    ora.build()
    # 1. fdrand 2D 10x10 (test_fdrand.jl sizes), seeded values, updateindex! flavour
    I, J, V = ora.fdrand_stream(10, 10, 1, seed=20240717)
    case_stream("fdrand_10x10", 100, 100, [(I, J, V, ora.UPDATE, True)])
    # 2. P1 FEM on a 4^3 Kuhn mesh, rawupdateindex!
    I, J, V = ora.fem_stream(4, 4, 4)
    case_stream("p1fem_4x4x4", 64, 64, [(I, J, V, ora.RAW, True)])
    # 3. mixed flavours, zeros, three splices (multi-splice flush of test_assembly.jl)
    rng = np.random.default_rng(42)
    batches = []
    for s in range(3):
        for b in range(4):
            cnt = 150
            I = rng.integers(1, 41, cnt)
            J = rng.integers(1, 31, cnt)
            V = rng.choice([0.0, -0.0, 1.0, -2.5, 0.125, 1e30], cnt) * rng.choice([1.0, 0.3], cnt)
            batches.append((I, J, V, int(rng.integers(0, 3)), b == 3))
    case_stream("mixed_40x30", 40, 30, batches)
    # 4. block reaction-diffusion 3x3x2 with 2 species
    I, J, V = ora.blockrd_stream(3, 3, 2, 2, seed=7)
    case_stream("blockrd_3x3x2x2", 36, 36, [(I, J, V, ora.UPDATE, True)])


if __name__ == "__main__":
    main()
