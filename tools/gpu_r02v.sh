#!/bin/bash
# Round 2, visit V (1 GPU): final state -- smoke, the whole GPU suite, and the convection kernel with the software-pipelined ascent
# loop against the previous build of the same file (isca_b200/lib/variants/conv_old)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 ) > gpurun_out/r02v_pytest_gpu.txt 2>&1
B="python bench.py --steps 96 --warmup 5 --no-extra --no-cpu-baseline --e2e-steps 1"
timeout 300 $B > gpurun_out/r02v_bench_new.json 2> /dev/null
ISCA_B200_LIB=$PWD/isca_b200/lib/variants/conv_old/libisca_b200.so timeout 300 $B > gpurun_out/r02v_bench_old.json 2> /dev/null
tail -3 gpurun_out/r02v_smoke.txt
tail -6 gpurun_out/r02v_pytest_gpu.txt
for v in new old; do python -c "
import json
d=json.load(open('gpurun_out/r02v_bench_$v.json'))
print('$v', 'ms_per_step', round(d['ms_per_step'],4), 'convection', d['kernel_groups_ms']['phys_convection'], 'vdd', d['kernel_groups_ms']['phys_vert_diff_down'])"; done
