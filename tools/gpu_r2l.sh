#!/bin/bash
# N=1 bench with all legs (validates the cfg1_small / reordered legs), twice for run-to-run spread
mkdir -p gpurun_out
for k in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/l_bench$k.json 2> gpurun_out/l_bench$k.err; echo "rc=$?" >> gpurun_out/l_bench$k.err
done
tail -n 4 gpurun_out/l_bench1.err; head -c 3000 gpurun_out/l_bench1.json
