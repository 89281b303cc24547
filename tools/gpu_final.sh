#!/bin/bash
# end-of-round evidence (development tool): smoke, GPU tests, ncu launch list + full captures of the three workloads,
# bench line and reference arm.  Outputs under gpurun_out/r2_*; the summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 19 -c 12 --csv --log-file gpurun_out/r2_launches_fem128.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-legs > gpurun_out/r2_launches_bench.log 2>&1
for w in fem128 fd200 rd96; do
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"grouped|runfold|run_bucket|run_totals|runpair|chunk_sort" -s 8 -c 5 -f -o gpurun_out/r2_full_$w \
    python tools/exp_stages.py $w > gpurun_out/r2_full_$w.log 2>&1
ncu -i gpurun_out/r2_full_$w.ncu-rep --page raw --csv > gpurun_out/r2_full_${w}_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/r2_full_${w}_raw.csv > gpurun_out/r2_ncu_full_$w.csv
rm -f gpurun_out/r2_full_${w}_raw.csv
done
# the generic insertion kernel (device triplets) and the values-only kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pack_grouped" -s 2 -c 1 -f -o gpurun_out/r2_full_pack \
    python tools/exp_pack.py > gpurun_out/r2_full_pack.log 2>&1
ncu -i gpurun_out/r2_full_pack.ncu-rep --page raw --csv > gpurun_out/r2_full_pack_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/r2_full_pack_raw.csv > gpurun_out/r2_ncu_full_pack_fem128.csv; rm -f gpurun_out/r2_full_pack_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reassemble_det" -s 3 -c 1 -f -o gpurun_out/r2_full_values \
    python tools/exp_values.py > gpurun_out/r2_full_values.log 2>&1
ncu -i gpurun_out/r2_full_values.ncu-rep --page raw --csv > gpurun_out/r2_full_values_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/r2_full_values_raw.csv > gpurun_out/r2_ncu_full_values_fd200.csv; rm -f gpurun_out/r2_full_values_raw.csv
timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/r2_stages.log 2>&1
timeout 300 python tools/exp_pack.py > gpurun_out/r2_pack.log 2>&1
timeout 300 python tools/exp_values.py > gpurun_out/r2_values.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref.json 2> gpurun_out/r2_ref.err; echo "ref rc=$?" >> gpurun_out/r2_ref.err
tail -n 2 gpurun_out/r2_smoke.log; tail -n 3 gpurun_out/r2_tests.log; cat gpurun_out/r2_stages.log; tail -n 2 gpurun_out/r2_bench.err; tail -n 2 gpurun_out/r2_ref.err; head -c 400 gpurun_out/r2_bench.json; echo; head -c 400 gpurun_out/r2_ref.json
