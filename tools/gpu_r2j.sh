#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.log
timeout 600 python tools/exp_stages.py fem128 fd200 rd96 > gpurun_out/j_stages.log 2>&1
XSB_DIST_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-legs > gpurun_out/j_n${N}_timing.json 2> gpurun_out/j_n${N}_timing.err; echo "rc=$?" >> gpurun_out/j_n${N}_timing.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-legs > gpurun_out/j_n$N.json 2> gpurun_out/j_n$N.err; echo "rc=$?" >> gpurun_out/j_n$N.err
tail -n 4 gpurun_out/j_tests.log; cat gpurun_out/j_stages.log; tail -n 3 gpurun_out/j_n${N}_timing.err; tail -n 3 gpurun_out/j_n$N.err
python - <<PY
import json
for f in ['gpurun_out/j_n${N}_timing.json','gpurun_out/j_n$N.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['roofline']['span_emission_to_csc']['ms'], d['config']['exchange'])
    except Exception as e:
        print(f, 'ERR', e)
PY
