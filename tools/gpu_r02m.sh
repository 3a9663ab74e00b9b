#!/bin/bash
# Round 2, visits M / N (4 and 8 GPUs): N-rank = 1-rank check at T170 L40, strong-scaling points of the dry core and of the MiMA headline
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $TR tools/multigpu_check.py T170 40 40 6 > gpurun_out/r02m_multigpu_check_t170_n$N.txt 2>&1
timeout 400 $TR bench.py --gpus $N --workload hs --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02m_bench_hs_n$N.json 2> gpurun_out/r02m_bench_hs_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 96 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02m_bench_mima_n$N.json 2> gpurun_out/r02m_bench_mima_n$N.err
if [ "$N" = "4" ]; then
  ( time timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rf --timeout 300 -s ) > gpurun_out/r02m_pytest_multirank_n$N.txt 2>&1
  tail -5 gpurun_out/r02m_pytest_multirank_n$N.txt
fi
du -sh gpurun_out
grep MULTIGPU gpurun_out/r02m_multigpu_check_t170_n$N.txt | cut -c1-900
tail -3 gpurun_out/r02m_bench_hs_n$N.err; cut -c1-700 gpurun_out/r02m_bench_hs_n$N.json
tail -3 gpurun_out/r02m_bench_mima_n$N.err; cut -c1-700 gpurun_out/r02m_bench_mima_n$N.json
