#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/h_n$N.json 2> gpurun_out/h_n$N.err; echo "n$N rc=$?" >> gpurun_out/h_n$N.err
tail -n 6 gpurun_out/h_n$N.err; head -c 1500 gpurun_out/h_n$N.json
