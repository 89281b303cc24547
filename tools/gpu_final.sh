#!/bin/bash
# end-of-round check (development tool): smoke, GPU tests, slab emit timing, bench line
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/z_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/z_tests.log
python tools/slabtime.py > gpurun_out/z_slab.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
tail -n 2 gpurun_out/z_smoke.log; tail -n 3 gpurun_out/z_tests.log; tail -n 2 gpurun_out/z_slab.log
