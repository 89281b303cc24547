#!/bin/bash
# N-GPU scaling session (development tool): FEM weak scaling + fdrand 400^3 strong scaling
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/n${N}_fem.json 2> gpurun_out/n${N}_fem.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 3 --warmup 3 --workload fd400 > gpurun_out/n${N}_fd400.json 2> gpurun_out/n${N}_fd400.err
tail -n 2 gpurun_out/n${N}_fem.err; tail -n 2 gpurun_out/n${N}_fd400.err
