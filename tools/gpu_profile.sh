#!/bin/bash
# ncu evidence for profiles/ (development tool): launch list of the bench command + one full capture of a step
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 33 -c 22 --csv --log-file gpurun_out/r1e_launches_fem128.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1e_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"group_count|group_scatter|colthread|emit_p1fem|onesweep|pair_" -s 26 -c 13 -f -o gpurun_out/r1e_full \
    python tools/exp_stages.py fem128 > gpurun_out/r1e_full.log 2>&1
ncu -i gpurun_out/r1e_full.ncu-rep --page raw --csv > gpurun_out/r1e_full_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/r1e_full_raw.csv > gpurun_out/r1e_ncu_full_fem128.csv
ls -la gpurun_out/r1e_*
python bench.py --steps 5 --warmup 3 > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1e_ref.json 2> gpurun_out/r1e_ref.err
