#!/bin/bash
# short multi-GPU visit: the N-rank run (HS + tracer, moist model) must reproduce the 1-rank run; bench line at N ranks
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multigpu_check.py T42 20 6 6 > gpurun_out/multi_t42_n$N.txt 2>&1
timeout 300 $TR tools/multigpu_check.py T170 40 4 4 > gpurun_out/multi_t170_n$N.txt 2>&1
ISCA_B200_NO_P2P=1 timeout 300 $TR tools/multigpu_check.py T42 20 6 6 > gpurun_out/multi_t42_nccl_n$N.txt 2>&1
timeout 400 $TR bench.py --gpus $N --steps 300 --warmup 10 --e2e-steps 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 400 python bench.py --steps 300 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.txt
grep -h MULTIGPU gpurun_out/multi_*_n$N.txt
tail -c 600 gpurun_out/bench_n$N.json
