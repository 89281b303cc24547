#!/bin/bash
mkdir -p gpurun_out
python tools/exp_values.py 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reassemble_det" -s 3 -c 1 \
  -o gpurun_out/p_values -f python tools/exp_values.py > gpurun_out/p_values.log 2>&1
ncu -i gpurun_out/p_values.ncu-rep --page raw --csv > gpurun_out/p_values_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/p_values_raw.csv > gpurun_out/p_values.csv
python tools/ncu_tables.py gpurun_out/p_values.csv | head -5
