#!/bin/bash
# one --set full capture of the FEM emitter (development tool)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"emit_p1fem_grouped" -s 2 -c 1 \
  -o gpurun_out/p_emit -f python tools/exp_stages.py fem128 > gpurun_out/p_emit.log 2>&1
ncu -i gpurun_out/p_emit.ncu-rep --page raw --csv > gpurun_out/p_emit_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/p_emit_raw.csv
