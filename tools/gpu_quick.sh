#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
python tools/devbench.py > gpurun_out/devbench_default.json 2> gpurun_out/devbench_default.err
python tools/devbench.py T85 40 > gpurun_out/devbench_t85.json 2>&1
python tools/physbench.py > gpurun_out/physbench.json 2>&1
