mkdir -p gpurun_out
for c in "" "16,768,768,768" "24,768,768,768" "32,768,768,768" "48,768,768,768" "24,96,768,768" "32,128,768,768" "96,768,768,768"; do
  if [ -z "$c" ]; then timeout 100 python scripts/probes/stack_time.py resnet50 768; else PDFUSION_B200_CHUNKS=$c timeout 100 python scripts/probes/stack_time.py resnet50 768; fi
done > gpurun_out/t3_chunks.txt 2>&1
cat gpurun_out/t3_chunks.txt
