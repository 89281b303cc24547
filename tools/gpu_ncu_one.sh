#!/bin/bash
# refresh the full capture of ONE workload: $1 = fem128 | fd200 | rd96
mkdir -p gpurun_out
w=$1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"grouped|runfold|run_bucket|run_totals|runpair|chunk_sort" -s 8 -c 5 -f -o gpurun_out/r2_full_$w \
    python tools/exp_stages.py $w > gpurun_out/r2_full_$w.log 2>&1
ncu -i gpurun_out/r2_full_$w.ncu-rep --page raw --csv > gpurun_out/r2_full_${w}_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/r2_full_${w}_raw.csv > gpurun_out/r2_ncu_full_$w.csv
rm -f gpurun_out/r2_full_${w}_raw.csv
python tools/ncu_tables.py gpurun_out/r2_ncu_full_$w.csv
python tools/exp_stages.py $w 2>&1 | tail -3
