#!/bin/bash
# same-box A/B of two builds of the library: scripts/ab_bench.sh <variant.so> [rounds]
# prints value / conv ms / preproc ms of the default build and the variant, alternating (box-to-box variance is ~4 %)
V=$1; R=${2:-2}
for r in $(seq $R); do
  for lib in default $V; do
    if [ $lib = default ]; then unset PDFUSION_B200_LIB; else export PDFUSION_B200_LIB=$PWD/$lib; fi
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', round(d['value']), 'step', round(d['ms_per_step'],3), 'conv', round(d['roofline']['ms'],3), 'pre', round(d['roofline_preproc']['ms'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
