#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/f_n2.json 2> gpurun_out/f_n2.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/f_n1.json 2> gpurun_out/f_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload fd400 > gpurun_out/f_n2_fd400.json 2> gpurun_out/f_n2_fd400.err
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/f_n2_tests.log 2>&1
tail -n 2 gpurun_out/f_n2_tests.log
