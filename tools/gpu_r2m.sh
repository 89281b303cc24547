#!/bin/bash
# phase times of the sharded step on every rank (diagnostic)
mkdir -p gpurun_out
N=${1:-4}
XSB_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-legs > gpurun_out/m_n$N.json 2> gpurun_out/m_n$N.err; echo "rc=$?" >> gpurun_out/m_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/h_n$N.json 2> gpurun_out/h_n$N.err; echo "rc=$?" >> gpurun_out/h_n$N.err
tail -n 3 gpurun_out/m_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/m_n$N.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['config']['exchange']))
d=json.loads(open('gpurun_out/h_n$N.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['cfg5']['ms_per_step'])
PY
