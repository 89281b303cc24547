#!/bin/bash
mkdir -p gpurun_out
python tools/exp_stages.py fem128 > gpurun_out/x_base.log 2>&1
XSB_NVCC_EXTRA="-DXSB_CT_D5=16" python extendablesparse.jl_b200/build.py --force > /dev/null 2>&1
python tools/exp_stages.py fem128 > gpurun_out/x_d16.log 2>&1
python -m pytest tests/test_gpu_parity.py -q -x -k "fem or precount" > gpurun_out/x_d16_tests.log 2>&1
tail -n 3 gpurun_out/x_base.log; tail -n 3 gpurun_out/x_d16.log; tail -n 2 gpurun_out/x_d16_tests.log
