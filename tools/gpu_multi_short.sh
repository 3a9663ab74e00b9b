#!/bin/bash
# short multi-GPU visit: N-rank T170 run must reproduce the 1-rank run; bench line at N ranks
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multigpu_check.py T170 40 4 > gpurun_out/multi_t170_n$N.txt 2>&1
timeout 300 $TR bench.py --gpus $N --steps 300 --warmup 10 --e2e-steps 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -h MULTIGPU gpurun_out/multi_t170_n$N.txt
tail -c 400 gpurun_out/bench_n$N.json
