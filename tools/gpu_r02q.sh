#!/bin/bash
# Round 2, visit Q (1 GPU): the GPU suite with the axisymmetric test-case options (free_atm_diff, do_sc_sst), smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02q_smoke.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 ) > gpurun_out/r02q_pytest_gpu.txt 2>&1
tail -3 gpurun_out/r02q_smoke.txt
tail -25 gpurun_out/r02q_pytest_gpu.txt
