#!/bin/bash
# Development tool: build differently tuned variants of the RRTMG kernels (tile size, CTAs per SM) as separate libraries under
# isca_b200/lib/variants/<name>/ (selected with ISCA_B200_LIB) so that one GPU visit can time them all:  tools/rrtm_variants.sh build | run
set -e
cd "$(dirname "$0")/.."
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -diag-suppress 550"
VARIANTS=("u2_4:-DISCA_LWC_MINB=4" "u2_5:-DISCA_LWC_MINB=5" "u2_6:-DISCA_LWC_MINB=6" "u4_3:-DISCA_LWC_UNROLL=4 -DISCA_LWC_MINB=3" "u4_4:-DISCA_LWC_UNROLL=4 -DISCA_LWC_MINB=4" "u4_5:-DISCA_LWC_UNROLL=4 -DISCA_LWC_MINB=5")
if [ "$1" = "build" ]; then
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}; defs=${v#*:}
    d=isca_b200/lib/variants/$name; mkdir -p $d
    $NVCC $FLAGS $defs -c isca_b200/csrc/rrtm.cu -o $d/rrtm.o
    objs=$(ls isca_b200/lib/*.o | grep -v "/rrtm.o")
    $NVCC -shared -o $d/libisca_b200.so $objs $d/rrtm.o -lcudart -ldl
    cuobjdump -res-usage $d/rrtm.o 2>/dev/null | grep -A1 "rrtmg_lw_col\|lw_record" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | tr '\n' ' '; echo " <- $name"
  done
else
  mkdir -p gpurun_out
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}
    echo "== $name"
    if [ "$name" = "u2_4" ]; then
      echo "-- g-point-per-thread LW kernel (ISCA_B200_RRTM_LW_GPOINT=1)"
      ISCA_B200_RRTM_LW_GPOINT=1 ISCA_B200_LIB=$PWD/isca_b200/lib/variants/$name/libisca_b200.so timeout 200 python tools/rrtm_bench.py 2>&1 | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('lw_ms', round(d['lw_kernel_ms'],2), 'sw_ms', round(d['sw_kernel_ms'],2), 'olr', repr(d['olr_mean']), 'sfc', repr(d['surf_lw_down_mean']), 'swdn', repr(d['sfc_sw_down_mean']), 'swup', repr(d['toa_sw_up_mean']))"
    fi
    ISCA_B200_LIB=$PWD/isca_b200/lib/variants/$name/libisca_b200.so timeout 200 python tools/rrtm_bench.py 2>&1 | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('lw_ms', round(d['lw_kernel_ms'],2), 'sw_ms', round(d['sw_kernel_ms'],2), 'olr', repr(d['olr_mean']), 'sfc', repr(d['surf_lw_down_mean']), 'swdn', repr(d['sfc_sw_down_mean']), 'swup', repr(d['toa_sw_up_mean']))"
  done
fi
