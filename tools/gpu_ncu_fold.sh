#!/bin/bash
# one --set full capture of the merge kernel (development tool): $1 = workload
mkdir -p gpurun_out
W=${1:-fem128}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"runfold" -s 2 -c 1 \
  -o gpurun_out/p_fold_$W -f python tools/exp_stages.py $W > gpurun_out/p_fold_$W.log 2>&1
ncu -i gpurun_out/p_fold_$W.ncu-rep --page raw --csv > gpurun_out/p_fold_${W}_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/p_fold_${W}_raw.csv | tail -1
