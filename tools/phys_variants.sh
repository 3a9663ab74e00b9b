#!/bin/bash
# Development tool: build the column-physics kernels with different resident-CTA requests (register caps) as separate libraries under
# isca_b200/lib/variants/<name>/ and time the MiMA T170 L40 step's kernel groups with each:  tools/phys_variants.sh build | run
set -e
cd "$(dirname "$0")/.."
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -diag-suppress 550 -Iinclude -Iisca_b200/csrc"
FILES="physics physics_conv physics_bm physics_turb physics_diff physics_surface"
VARIANTS=("p_old:-DISCA_COL_MINB=1 -DISCA_VDD_MINB=1 -DISCA_SF_MINB=1" "p_vdd5:-DISCA_VDD_MINB=5" "p_vdd7:-DISCA_VDD_MINB=7" "p_sf7:-DISCA_SF_MINB=7")
if [ "$1" = "build" ]; then
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}; defs=${v#*:}
    d=isca_b200/lib/variants/$name; mkdir -p $d
    objs=""
    for f in $FILES; do $NVCC $FLAGS $defs -c isca_b200/csrc/$f.cu -o $d/$f.o; objs="$objs $d/$f.o"; done
    rest=$(ls isca_b200/lib/*.o | grep -v "/physics\(_conv\|_bm\|_turb\|_diff\|_surface\)\?\.o")
    $NVCC -shared -o $d/libisca_b200.so $rest $objs -lcudart -ldl
    echo "built $name"
  done
else
  mkdir -p gpurun_out
  for name in default p_old p_vdd5 p_vdd7 p_sf7; do
    echo "== $name"
    if [ "$name" = "default" ]; then unset ISCA_B200_LIB; else export ISCA_B200_LIB=$PWD/isca_b200/lib/variants/$name/libisca_b200.so; fi
    timeout 400 python bench.py --steps 48 --warmup 3 --spinup 384 --no-extra --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
g=d['kernel_groups_ms']
print('ms_per_step', round(d['ms_per_step'],4), {k: round(v,4) for k,v in g.items() if k.startswith('phys_')})"
  done
fi
