#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reassembl or frozen or values" > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log
for w in fem128 fd200 rd96; do
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"grouped|runfold|run_bucket|run_totals|runpair|chunk_sort" -s 8 -c 5 -f -o gpurun_out/g_full_$w \
    python tools/exp_stages.py $w > gpurun_out/g_full_$w.log 2>&1
ncu -i gpurun_out/g_full_$w.ncu-rep --page raw --csv > gpurun_out/g_full_${w}_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/g_full_${w}_raw.csv > gpurun_out/g_ncu_$w.csv
done
ncu -i gpurun_out/g_full_fem128.ncu-rep --page source --csv -k regex:runfold > gpurun_out/g_fold_source.csv 2>/dev/null
python - <<'PY' > gpurun_out/g_vo.log 2>&1
import sys; sys.path.insert(0,'.')
import __graft_entry__ as ge; ge.build()
import xsparse_b200 as xsb, torch, ctypes as C, argparse, bench
a=argparse.Namespace(vo_n=200, steps=5)
print(bench.measure_values_only(a, xsb, torch, 6543.4, 0))
PY
tail -3 gpurun_out/g_tests.log; cat gpurun_out/g_vo.log | tail -3
