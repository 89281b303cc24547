#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/f_n2.json 2> gpurun_out/f_n2.err; echo "n2 rc=$?" >> gpurun_out/f_n2.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/f_n1.json 2> gpurun_out/f_n1.err; echo "n1 rc=$?" >> gpurun_out/f_n1.err
tail -n 8 gpurun_out/f_tests.log; tail -n 12 gpurun_out/f_n2.err; head -c 2500 gpurun_out/f_n2.json; echo; tail -n 3 gpurun_out/f_n1.err; head -c 600 gpurun_out/f_n1.json
