#!/bin/bash
# N ranks: bench with --verify over the peer transport, then over NCCL send/receive
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 --verify > gpurun_out/n_peer_n$N.json 2> gpurun_out/n_peer_n$N.err; echo "rc=$?" >> gpurun_out/n_peer_n$N.err
XSB_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $N --steps 5 --warmup 3 --verify --no-legs > gpurun_out/n_nccl_n$N.json 2> gpurun_out/n_nccl_n$N.err; echo "rc=$?" >> gpurun_out/n_nccl_n$N.err
tail -n 3 gpurun_out/n_peer_n$N.err gpurun_out/n_nccl_n$N.err; python - <<PY
import json
for t in ("peer","nccl"):
    d=json.loads(open(f'gpurun_out/n_{t}_n$N.json').read().strip().splitlines()[-1])
    print(t, d['ms_per_step'], d['config']['exchange'].get('transport'), d.get('parity_checked',{}).get('slabs_identical_to_single_gpu_assembly'), (d.get('cfg5') or {}).get('ms_per_step'))
PY
