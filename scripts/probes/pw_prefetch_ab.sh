mkdir -p gpurun_out
timeout 120 python scripts/probes/op_times.py resnet50 768 > gpurun_out/t31_op_pf0.txt 2>&1
PDFUSION_B200_PW_PREFETCH=1 timeout 120 python scripts/probes/op_times.py resnet50 768 > gpurun_out/t31_op_pf1.txt 2>&1
tail -n 1 gpurun_out/t31_op_pf0.txt
tail -n 1 gpurun_out/t31_op_pf1.txt
