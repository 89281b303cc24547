#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c_tests.log
timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/c_stages.log 2>&1; echo "stages rc=$?" >> gpurun_out/c_stages.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -q -x > gpurun_out/c_fullsize.log 2>&1; echo "fullsize rc=$?" >> gpurun_out/c_fullsize.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?" >> gpurun_out/c_bench.err
tail -n 8 gpurun_out/c_tests.log; cat gpurun_out/c_stages.log; tail -n 8 gpurun_out/c_fullsize.log; tail -n 5 gpurun_out/c_bench.err; head -c 3000 gpurun_out/c_bench.json
