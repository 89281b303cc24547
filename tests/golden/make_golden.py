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


def digest(*arrays):
    import hashlib

    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


# Streams large enough for the product path of the flush (grouping by column + thread fold; >= 32768 records):
# too big to commit as arrays, so only SHA-256 digests of the oracle's result are recorded.  The CPU test
# regenerates them with the oracle (drift guard), the GPU test must reach the same bytes through the
# library's own on-device emitters.
HASHED = {
    "p1fem_20": dict(kind="fem", dims=(20, 20, 20), flavour="raw"),
    "fdrand_40": dict(kind="fd", dims=(40, 40, 40), seed=20240717, flavour="update"),
    "blockrd_12x10x8x4": dict(kind="rd", dims=(12, 10, 8, 4), seed=3, flavour="update"),
}


def hashed_case(spec):
    k, dims = spec["kind"], spec["dims"]
    if k == "fem":
        I, J, V = ora.fem_stream(*dims)
        n = dims[0] * dims[1] * dims[2]
    elif k == "fd":
        I, J, V = ora.fdrand_stream(*dims, seed=spec["seed"])
        n = dims[0] * dims[1] * dims[2]
    else:
        I, J, V = ora.blockrd_stream(*dims, seed=spec["seed"])
        n = dims[0] * dims[1] * dims[2] * dims[3]
    A = ora.OracleExt(n, n)
    A.insert_batch(I, J, V, ora.RAW if spec["flavour"] == "raw" else ora.UPDATE)
    cp, rv, nz = A.csc()
    out = {"n": n, "records": int(len(V)), "nnz": int(len(nz)), "csc": digest(cp, rv, nz)}
    bcp, brv, bl = A.pointblock(4)
    out["pointblock4"] = digest(bcp, brv, bl)
    out["nnz_blocks"] = int(len(brv))
    return out


def main_hashes():
    import json

    ora.build()
    res = {name: dict(spec, **hashed_case(spec)) for name, spec in HASHED.items()}
    with open(os.path.join(HERE, "hashes.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
    main_hashes()
