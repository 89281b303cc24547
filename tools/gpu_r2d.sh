#!/bin/bash
mkdir -p gpurun_out
python tools/diag_splice.py > gpurun_out/d_diag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"runfold" -s 2 -c 1 -f -o gpurun_out/d_fold \
    python tools/exp_stages.py fem128 > gpurun_out/d_fold.log 2>&1
ncu -i gpurun_out/d_fold.ncu-rep --page source --csv > gpurun_out/d_fold_source.csv 2>/dev/null
ncu -i gpurun_out/d_fold.ncu-rep --page raw --csv > gpurun_out/d_fold_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"p1fem_grouped" -s 2 -c 1 -f -o gpurun_out/d_emit \
    python tools/exp_stages.py fem128 > gpurun_out/d_emit.log 2>&1
ncu -i gpurun_out/d_emit.ncu-rep --page source --csv > gpurun_out/d_emit_source.csv 2>/dev/null
ncu -i gpurun_out/d_emit.ncu-rep --page raw --csv > gpurun_out/d_emit_raw.csv 2>/dev/null
cat gpurun_out/d_diag.log
