for m in 0 1 2 3; do
  echo "== CONV_PROBE=$m"
  PDFUSION_B200_CONV_PROBE=$m timeout 200 python scripts/probes/op_times.py 2>&1 | grep -E "op  [5-9]|op 1[05]|sum"
done
