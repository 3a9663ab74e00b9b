#!/bin/bash
# Round 2, visit O (N GPUs): bench lines after the clock-sampler fix (one nvidia-smi on rank 0, started before the timed region)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus $N --workload hs --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02o_bench_hs_n$N.json 2> gpurun_out/r02o_bench_hs_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 96 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02o_bench_mima_n$N.json 2> gpurun_out/r02o_bench_mima_n$N.err
for w in hs mima; do tail -2 gpurun_out/r02o_bench_${w}_n$N.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r02o_bench_${w}_n$N.json'))
print('$w N=$N ms_per_step', d['ms_per_step'], 'wall', d['config'].get('wall_ms_per_step'), 'value', d['value'], 'clocks', d['clocks'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"; done
