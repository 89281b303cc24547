#!/bin/bash
# one-off GPU session: new-feature tests, bench, tuning experiments (development tool)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
nproc >> gpurun_out/c1_gpu.txt; free -g >> gpurun_out/c1_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "triplet or micro or fem or bounds" > gpurun_out/c1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c1_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c1_ref.json 2> gpurun_out/c1_ref.err
timeout 300 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/c1_exp_base.log 2>&1
for v in "-DXSB_GP_W=1024" "-DXSB_GP_W=2048" "-DXSB_CT_U=3" "-DXSB_CT_U=4" "-DXSB_CT_U=1"; do
  tag=$(echo "$v" | tr -d '=-' )
  XSB_NVCC_EXTRA="$v" python extendablesparse.jl_b200/build.py --force > /dev/null 2>&1
  timeout 300 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/c1_exp_$tag.log 2>&1
done
echo done
