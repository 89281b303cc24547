#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pack_grouped" -s 2 -c 1 \
  -o gpurun_out/p_pack -f python tools/exp_pack.py > gpurun_out/p_pack.log 2>&1
ncu -i gpurun_out/p_pack.ncu-rep --page raw --csv > gpurun_out/p_pack_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/p_pack_raw.csv > gpurun_out/p_pack.csv
python tools/ncu_tables.py gpurun_out/p_pack.csv
