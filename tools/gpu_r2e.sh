#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/e_tests.log
timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/e_stages.log 2>&1; echo "stages rc=$?" >> gpurun_out/e_stages.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -q > gpurun_out/e_fullsize.log 2>&1; echo "fullsize rc=$?" >> gpurun_out/e_fullsize.log
tail -n 8 gpurun_out/e_tests.log; cat gpurun_out/e_stages.log; tail -n 4 gpurun_out/e_fullsize.log
