for lib in default variant_fix1.so; do
  if [ $lib = default ]; then unset PDFUSION_B200_LIB; else export PDFUSION_B200_LIB=$PWD/$lib; fi
  for prio in none conv pre; do
    export PDFUSION_B200_PRIO=$prio
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib prio=$prio', round(d['value']), 'step', round(d['ms_per_step'],3), 'conv', round(d['roofline']['ms'],3), 'pre', round(d['roofline_preproc']['ms'],3))"
  done
  unset PDFUSION_B200_PRIO
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-overlap 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib no-overlap', round(d['value']), 'step', round(d['ms_per_step'],3), 'conv', round(d['roofline']['ms'],3), 'pre', round(d['roofline_preproc']['ms'],3))"
done
