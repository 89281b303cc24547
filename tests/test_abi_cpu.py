"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
and its host-side stream arithmetic agrees with the oracle.  No compute call needs a GPU here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def xsb():
    import __graft_entry__ as ge

    ge.build()
    import xsparse_b200

    return xsparse_b200


def header_functions():
    text = open(os.path.join(ROOT, "include", "xsparse_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xsb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(xsb):
    names = header_functions()
    assert len(names) >= 35
    lib = ctypes.CDLL(xsb.capi.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/xsparse_b200.h but not exported"
    # and the Python binding covers exactly the declared surface
    assert sorted(xsb.capi.SIGNATURES) == names


def test_version(xsb):
    assert xsb.capi.lib().xsb_version() == 100


def test_stream_counts_match_oracle(xsb, oracle):
    for dims in [(1, 1, 1), (2, 1, 1), (3, 1, 1), (100, 1, 1), (10, 10, 1), (2, 2, 1), (3, 3, 1), (5, 5, 5), (7, 6, 5),
                 (2, 3, 4), (1, 5, 1), (1, 1, 7), (3, 1, 4), (100, 100, 1), (33, 17, 9)]:
        assert xsb.capi.stream_count_fdrand(*dims) == oracle.fdrand_count(*dims), dims
    assert xsb.capi.stream_count_p1fem(128, 128, 128) == oracle.fem_count(128, 128, 128)
    assert xsb.capi.stream_count_blockrd(96, 96, 96, 4) == oracle.blockrd_count(96, 96, 96, 4)
    assert xsb.capi.stream_count_blockrd(3, 4, 5, 2) == oracle.blockrd_count(3, 4, 5, 2)


def test_no_cpu_fallback(xsb):
    """Without a CUDA device the product fails loudly instead of computing on the host."""
    if xsb.capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(xsb.XsbError) as e:
        xsb.Handle(10, 10)
    assert e.value.code == xsb.capi.ECUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "extendablesparse.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "xsb_oracle", "libxsb_oracle", "ora_"):
                    assert needle not in src, f"{f} references the oracle ({needle})"


def test_reference_arm_prints_one_json_line():
    """bench.py --impl reference: the oracle port of BOTH reference paths (serial and partitioned) timed on the
    host cores, one JSON line with the contract's keys; runs without a GPU."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-mesh", "12",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "entries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["serial_value"] > 0 and cb["cores"] >= 1
    assert cb["value"] == max(cb["serial_value"], cb["mt_value"] or 0.0)


def test_roofline_accounting_covers_every_flush_path():
    """bench.roofline(): the dominant stage, its algorithmic bytes and the whole-flush fraction for the product
    path (bucketed or sorted pairs, with and without counting at insertion) and the fallback paths."""
    import bench

    ms = {"ms_group_count": 0.66, "ms_group_scatter": 1.48, "ms_pair_sort": 0.27, "ms_fold": 1.18, "ms_compact": 0.0,
          "ms_sort": 1.75, "ms_reduce": 1.18, "ms_total": 3.66}
    st = {"n_inserted": 245805960, "nnz_old": 0, "nnz_new": 31065598, "group_pairs": 11735358, "column_path": 3,
          "sort_passes": 0, "precounted": 1.0, "direct_fold": 1}
    r = bench.roofline(st, ms, 2097152, 6543.4, "measured", None)
    assert r["kernel"].startswith("group_scatter_kernel") and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["achieved"] - (32 * 245805960 + 8 * 11735358) / 1.48e-3 / 1e9) < 1e-6
    assert abs(r["frac"] - r["achieved"] / 6543.4) < 1e-12 and r["traffic"] == 4.0631e9 + 3.8998e9
    assert abs(r["flush"]["algorithmic_bytes"] - bench.flush_bytes(245805960, 0, 31065598, 2097152)) < 1
    assert any(k.startswith("pair buckets") for k in r["kernels"])
    st2 = dict(st, sort_passes=3, precounted=0.0)
    r2 = bench.roofline(st2, dict(ms, ms_pair_sort=2.0), 2097152, 6543.4, "measured", None)
    assert r2["kernel"].startswith("onesweep_kernel on the (column, chunk) pairs") and r2["launches_per_step"] == 3
    for path in (0, 2):
        st3 = dict(st, column_path=path, sort_passes=6, group_pairs=0)
        r3 = bench.roofline(st3, dict(ms, ms_sort=14.0), 2097152, 6543.4, "measured", None)
        assert r3["kernel"].startswith("onesweep_kernel (one radix pass") and r3["traffic"] is None
