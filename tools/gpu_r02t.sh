#!/bin/bash
# Round 2, visit T (1 GPU): the derived-diagnostics test after the tolerance fix + the parity file it lives in
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --timeout 600 ) > gpurun_out/r02t_pytest_parity.txt 2>&1
tail -8 gpurun_out/r02t_pytest_parity.txt
