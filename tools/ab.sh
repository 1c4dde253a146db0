#!/bin/bash
# A/B builds of libclane_b200.so in ONE gpurun call (same box, back to back): the in-tree build vs variants
# compiled with extra -D flags.  Usage (from the repo root, in the build container):
#   tools/ab.sh [--tests] "-DCLANE_ROW_WARPS_PER_SM=28" "-DCLANE_ROW_WARPS=1" ...
# Prints ms/step and the row-kernel time of every variant, twice each.  With --tests the GPU parity tests run first on
# the in-tree build.
set -e
cd "$(dirname "$0")/.."
TESTS=0
if [ "$1" == "--tests" ]; then TESTS=1; shift; fi
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-O2,-fvisibility=hidden -shared -cudart static -I include -I clane_b200/csrc"
python -m clane_b200.build
LIBS="''"
i=0
for v in "$@"; do
  i=$((i + 1))
  nvcc $FLAGS $v -o tools/_ab_$i.so clane_b200/csrc/*.cu
  LIBS="$LIBS tools/_ab_$i.so"
done
T=""
if [ $TESTS == 1 ]; then T="timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4;"; fi
/usr/local/graft/bin/gpurun --timeout 1500 -- "$T for v in $LIBS; do for k in 1 2; do CLANE_LIB=\$v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-converge 2>/dev/null | python -c \"import json,sys; j=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); r=j['roofline']; print('\$v', 'ms/step', round(j['ms_per_step'],5), 'rows_ms', round(r['kernel_ms'],5), 'tail_ms', round(r['l1_tail_ms'],5), 'frac', round(r['frac'],3))\"; done; done" 2>&1 | grep "passed\|failed\|ms/step\|status\|Error\|error"
rm -f tools/_ab_*.so
