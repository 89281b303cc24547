#!/bin/bash
# final multi-GPU line: bench.py --gpus N --verify (peer exchange)
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --verify > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2_bench_n$N.err
tail -n 2 gpurun_out/r2_bench_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value']/1e9, d['config']['exchange'].get('transport'), d.get('parity_checked',{}).get('slabs_identical_to_single_gpu_assembly'), (d.get('cfg5') or {}).get('ms_per_step'), d['e2e']['value']/1e9, d['roofline']['frac'])
PY
