#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/k_stages.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/k_tests.log
cat gpurun_out/k_stages.log; tail -3 gpurun_out/k_tests.log
