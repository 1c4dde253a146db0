#!/bin/bash
# One gpurun call: bench.py (arxiv shape, N=1, no CPU arm, no convergence run) once per environment setting given
# as an argument ("" = defaults), e.g.   tools/meas.sh "" "CLANE_SWEEP_BATCH=30" "CLANE_NO_GRAPHS=1"
cd "$(dirname "$0")/.."
python -m clane_b200.build >/dev/null
CMD=""
for v in "$@"; do
  CMD="$CMD echo \"[$v]\"; env $v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-converge 2>/dev/null | python -c \"import json,sys; j=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); r=j['roofline']; print('ms/step', round(j['ms_per_step'],5), 'rows_ms', round(r['kernel_ms'],5), 'tail_ms', round(r['l1_tail_ms'],5), 'frac', round(r['frac'],3), 'step_frac', round(r['whole_step_gbs']/r['peak'],3), 'e2e_s', round(j['e2e']['seconds'],4), 'build_p_ms', round(j['build_p']['ms'],4), j['last_amount'])\";"
done
/usr/local/graft/bin/gpurun --timeout 1200 -- "$CMD" 2>&1 | grep -v "^\[gpurun\] sending"
