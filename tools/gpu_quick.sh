#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
for p in 1 2 3 4; do
  ISCA_B200_PIPE=$p python tools/devbench.py > gpurun_out/devbench_pipe$p.json 2> gpurun_out/devbench_pipe$p.err
done
ISCA_B200_PIPE=3 python tools/devbench.py T85 40 > gpurun_out/devbench_t85.json 2>&1
