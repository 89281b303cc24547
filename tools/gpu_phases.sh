#!/bin/bash
# per-rank device phases of the sharded FEM step (diagnostic; the events synchronise every step)
mkdir -p gpurun_out
N=${1:-2}
XSB_DIST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 --no-legs --e2e-mesh 32 > gpurun_out/q_n$N.json 2> gpurun_out/q_n$N.err; echo "rc=$?" >> gpurun_out/q_n$N.err
tail -n 2 gpurun_out/q_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/q_n$N.json').read().strip().splitlines()[-1])
print(d['ms_per_step'])
for r,p in enumerate(d['config']['exchange']['device_phase_ms_last_step_per_rank']): print(r,p)
print(json.dumps(d['roofline']['stage_ms_per_step']))
PY
