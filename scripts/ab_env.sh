#!/bin/bash
# same-box A/B over values of one environment variable: scripts/ab_env.sh VAR "v1 v2 ..." [rounds] [extra bench args]
# (an empty value "-" means: variable unset)
VAR=$1; VALS=$2; R=${3:-2}; shift 3
for r in $(seq $R); do
  for v in $VALS; do
    if [ "$v" = "-" ]; then unset $VAR; else export $VAR=$v; fi
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$VAR=$v', round(d['value']), 'step', round(d['ms_per_step'],3), 'conv', round(d['roofline']['ms'],3), 'pre', round(d['roofline_preproc']['ms'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
