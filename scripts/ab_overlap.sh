#!/bin/bash
# same-box A/B: two-stream overlap vs one stream.  scripts/ab_overlap.sh [rounds]
R=${1:-2}
for r in $(seq $R); do
  for mode in overlap no-overlap; do
    if [ $mode = overlap ]; then A="--overlap"; else A="--no-overlap"; fi
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline $A 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$mode', round(d['value']), 'step', round(d['ms_per_step'],3), 'conv', round(d['roofline']['ms'],3), 'pre', round(d['roofline_preproc']['ms'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
