#!/bin/bash
# N ranks: peer transport with the exchange started early (interface first) vs at the end of the emission
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --verify > gpurun_out/o_early_n$N.json 2> gpurun_out/o_early_n$N.err; echo "rc=$?" >> gpurun_out/o_early_n$N.err
XSB_EXCHANGE_EARLY=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/o_late_n$N.json 2> gpurun_out/o_late_n$N.err; echo "rc=$?" >> gpurun_out/o_late_n$N.err
tail -n 2 gpurun_out/o_early_n$N.err gpurun_out/o_late_n$N.err; python - <<PY
import json
for t in ("early","late"):
    d=json.loads(open(f'gpurun_out/o_{t}_n$N.json').read().strip().splitlines()[-1])
    print(t, d['ms_per_step'], d['config']['exchange'].get('transport'), d.get('parity_checked',{}).get('slabs_identical_to_single_gpu_assembly'), (d.get('cfg5') or {}).get('ms_per_step'))
PY
