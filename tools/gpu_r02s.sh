#!/bin/bash
# Round 2, visit S (1 GPU): the GPU suite with the derived spectral_diagnostics fields
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 ) > gpurun_out/r02s_pytest_gpu.txt 2>&1
tail -25 gpurun_out/r02s_pytest_gpu.txt
