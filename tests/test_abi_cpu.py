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


def test_julia_glue_only_calls_declared_entry_points():
    """The Julia glue cannot be executed in this image: at least every symbol it ccalls must exist in the header,
    and the entry points of the drop-in sequence must be among them."""
    glue = open(os.path.join(ROOT, "extendablesparse.jl_b200", "julia", "ExtendableSparseB200.jl")).read()
    called = set(re.findall(r"ccall\(\(:(xsb_[a-z0-9_]+)", glue))
    declared = set(header_functions())
    assert called and called <= declared, sorted(called - declared)
    for name in ("xsb_create", "xsb_set_csc", "xsb_insert_triplets", "xsb_flush", "xsb_fetch_csc", "xsb_destroy",
                 "xsb_create_slab", "xsb_peer_exchange_create", "xsb_route_pack_peer", "xsb_route_unpack_peer"):
        assert name in called, name


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
    """bench.roofline(): roofline.frac is SURVEY.md 8(d)'s fraction -- the flush's algorithmic bytes over the device
    time of the whole flush! -- for the product path and the fallback paths; per-kernel detail beside it."""
    import bench

    ms = {"ms_group_count": 0.0, "ms_pair_sort": 0.17, "ms_fold": 1.79, "ms_sort": 0.17, "ms_reduce": 1.79, "ms_total": 1.99}
    st = {"n_inserted": 245805960, "nnz_old": 0, "nnz_new": 31065598, "group_pairs": 10679412, "column_path": 4,
          "sort_passes": 0, "precounted": 1.0, "direct_fold": 1}
    b = bench.flush_bytes(245805960, 0, 31065598, 2097152)
    assert b == 16 * 245805960 + 16 * 31065598 + 16 * (2097152 + 1)
    r = bench.roofline(st, ms, 1.9, 2097152, 6543.4, "measured", None)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["bytes_per_launch"] == b and r["ms_per_launch"] == 1.99
    assert abs(r["achieved"] - b / 1.99e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / 6543.4) < 1e-12
    assert r["traffic"] is None and "runfold_kernel" in r["kernel"]
    assert abs(r["span_emission_to_csc"]["ms"] - 3.89) < 1e-9
    assert any(k.startswith("run buckets") for k in r["kernels"])
    for path in (0, 2, 3):
        st3 = dict(st, column_path=path, sort_passes=6, group_pairs=0)
        r3 = bench.roofline(st3, dict(ms, ms_sort=14.0, ms_total=18.0), None, 2097152, 6543.4, "measured", "fem128")
        assert abs(r3["achieved"] - b / 18.0e-3 / 1e9) < 1e-6 and r3["traffic"] is None
    assert bench.values_only_bytes(95760000, 55760000) == 12 * 95760000 + 8 * 55760000
