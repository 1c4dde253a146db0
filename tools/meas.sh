#!/bin/bash
# One gpurun call: bench.py (arxiv shape, N=1, no CPU arm, no convergence run) once per environment setting given
# as an argument (NAME=VALUE; use X=1 for the defaults), e.g.   tools/meas.sh X=1 CLANE_SWEEP_BATCH=30 CLANE_NO_GRAPHS=1
cd "$(dirname "$0")/.."
python -m clane_b200.build >/dev/null
CMD=""
for v in "$@"; do
  CMD="$CMD echo \"[$v]\"; env $v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-converge 2>/dev/null | python tools/pick.py;"
done
/usr/local/graft/bin/gpurun --timeout 1200 -- "$CMD" 2>&1 | grep -v "^\[gpurun\] sending"
