#!/usr/bin/env python
"""Probe for a Julia runtime + ExtendableSparse.jl and, when present, run the REAL reference.

    python tools/julia_ref.py probe                 -> prints what was found (exit 0 either way)
    python tools/julia_ref.py dump  [outdir]        -> the package's own CSCs for the golden streams of tests/golden/
                                                       (written next to the oracle's: tests/golden/julia_<case>.npz via .csv)
    python tools/julia_ref.py time  <mesh> <reps>   -> seconds per FEM assembly with the package (serial ExtendableSparseMatrix)

Nothing here can run in the build container (no `julia`, no network); bench.py --impl reference calls probe() and
falls back to the oracle port when it returns None.  The Julia side is tools/julia_ref.jl.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "tools", "julia_ref.jl")


def probe():
    """Path of a julia binary that can `using ExtendableSparse`, or None."""
    cands = [shutil.which("julia"), os.path.join(ROOT, "baseline", "_ref", "bin", "julia"),
             os.path.join(ROOT, "baseline", "_ref", "julia", "bin", "julia")]
    for exe in cands:
        if not exe or not os.path.exists(exe):
            continue
        env = dict(os.environ)
        depot = os.path.join(ROOT, "baseline", "_ref", "depot")
        if os.path.isdir(depot):
            env["JULIA_DEPOT_PATH"] = depot
        try:
            r = subprocess.run([exe, "--startup-file=no", "-e", "using ExtendableSparse; print(pkgversion(ExtendableSparse))"],
                               capture_output=True, text=True, timeout=600, env=env)
        except Exception:  # noqa: BLE001
            continue
        if r.returncode == 0:
            return exe, r.stdout.strip(), env
    return None


def run(args):
    found = probe()
    if found is None:
        return None
    exe, _, env = found
    return subprocess.run([exe, "--startup-file=no", "-t", "1", JL, *map(str, args)], capture_output=True, text=True,
                          timeout=3600, env=env)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "probe"
    if cmd == "probe":
        f = probe()
        print("julia + ExtendableSparse: " + (f"{f[0]} (ExtendableSparse {f[1]})" if f else "not found"))
    else:
        r = run(sys.argv[1:])
        if r is None:
            print("julia + ExtendableSparse not found: nothing done")
        else:
            sys.stdout.write(r.stdout)
            sys.stderr.write(r.stderr)
            sys.exit(r.returncode)
