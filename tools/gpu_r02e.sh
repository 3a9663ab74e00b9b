#!/bin/bash
# Round 2, visit E (2 GPUs): N-rank = 1-rank parity (peer-memory and NCCL transports), strong-scaling points of the headline and of the dry core
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rf --timeout 600 -s ) > gpurun_out/r02e_pytest_multirank.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tools/multigpu_check.py T170 40 40 6 > gpurun_out/r02e_multigpu_check_t170_n2.txt 2>&1
timeout 600 $TR bench.py --gpus 2 --workload hs --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02e_bench_hs_n2.json 2> gpurun_out/r02e_bench_hs_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 96 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_mima_n2.json 2> gpurun_out/r02e_bench_mima_n2.err
du -sh gpurun_out
tail -15 gpurun_out/r02e_pytest_multirank.txt
grep MULTIGPU gpurun_out/r02e_multigpu_check_t170_n2.txt | cut -c1-900
tail -3 gpurun_out/r02e_bench_hs_n2.err; cut -c1-1500 gpurun_out/r02e_bench_hs_n2.json
tail -3 gpurun_out/r02e_bench_mima_n2.err; cut -c1-1500 gpurun_out/r02e_bench_mima_n2.json
