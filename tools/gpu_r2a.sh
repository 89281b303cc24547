#!/bin/bash
# round-2 development run: smoke, GPU tests, a memcheck pass over small cases, stage timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "test_emit_p1fem or test_micro or test_assembly_random" > gpurun_out/a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/a_memcheck.log
timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/a_stages.log 2>&1; echo "stages rc=$?" >> gpurun_out/a_stages.log
XSB_PRECOUNT=0 timeout 600 python tools/exp_stages.py fem128 fd200 > gpurun_out/a_stages_noprecount.log 2>&1
XSB_RUNS=0 timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/a_stages_oldpath.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x > gpurun_out/a_fullsize.log 2>&1; echo "fullsize rc=$?" >> gpurun_out/a_fullsize.log
tail -n 3 gpurun_out/a_smoke.log; tail -n 15 gpurun_out/a_tests.log; tail -n 5 gpurun_out/a_memcheck.log; cat gpurun_out/a_stages.log; tail -n 5 gpurun_out/a_fullsize.log
