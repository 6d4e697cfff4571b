#!/bin/bash
# tuning helper (GPU box): conv-stack time of the C2 workload for several depth-first chunk sizes
for c in "768,768,768,768" "56" "62" "68" "34" "45" "56,768,768,768" "68,768,768,768" "62,124,768,768" "56,224,768,768"; do
  PDFUSION_B200_CHUNKS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$c', round(d['value']), round(d['ms_per_step'],3), 'conv_ms', round(d['roofline']['ms'],3), 'frac', round(d['roofline']['frac'],3), 'pre_ms', round(d['roofline_preproc']['ms'],3))"
done
