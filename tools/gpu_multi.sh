#!/bin/bash
# multi-GPU visit: N-rank run must reproduce the 1-rank run; timing at T170 L40
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tools/multigpu_check.py T42 20 6 > gpurun_out/multi_check_n$N.txt 2>&1
$TR tools/multigpu_check.py T170 40 4 > gpurun_out/multi_t170_n$N.txt 2>&1
ISCA_B200_GRAPH_MULTI=1 $TR tools/multigpu_check.py T170 40 4 > gpurun_out/multi_t170_graph_n$N.txt 2>&1
$TR bench.py --gpus $N --steps 300 --warmup 10 --e2e-steps 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -h MULTIGPU gpurun_out/multi_*_n$N.txt
tail -c 600 gpurun_out/bench_n$N.json
