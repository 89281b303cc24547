#!/bin/bash
# round-2 development run: GPU tests, stage timings, ncu full capture of the grouped-chunk path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b_tests.log
timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/b_stages.log 2>&1; echo "stages rc=$?" >> gpurun_out/b_stages.log
for w in fem128 fd200; do
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"grouped|runfold|run_bucket|run_totals|runpair|chunk_sort" -s 8 -c 5 -f -o gpurun_out/b_full_$w \
    python tools/exp_stages.py $w > gpurun_out/b_full_$w.log 2>&1
ncu -i gpurun_out/b_full_$w.ncu-rep --page raw --csv > gpurun_out/b_full_${w}_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/b_full_${w}_raw.csv > gpurun_out/b_ncu_$w.csv
done
tail -n 8 gpurun_out/b_tests.log; cat gpurun_out/b_stages.log
