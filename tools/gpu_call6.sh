#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_tests.log 2>&1
tail -15 gpurun_out/c6_tests.log
timeout 300 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/c6_exp_on.log 2>&1
XSB_PRECOUNT=0 timeout 300 python tools/exp_stages.py fem128 > gpurun_out/c6_exp_off.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
cat gpurun_out/c6_exp_on.log gpurun_out/c6_exp_off.log
