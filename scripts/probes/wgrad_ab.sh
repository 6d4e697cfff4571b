mkdir -p gpurun_out
timeout 100 python scripts/probes/wgrad_shapes.py > gpurun_out/t24_wgrad_shapes.txt 2>&1; cat gpurun_out/t24_wgrad_shapes.txt
for w in 1 2 3 4; do PDFUSION_B200_WGRAD_WAVES=$w FT_STEP_REPS=10 timeout 100 python scripts/probes/ft_step.py 4 64 3; done > gpurun_out/t24_waves.txt 2>&1; cat gpurun_out/t24_waves.txt
(timeout 300 python -m pytest tests/test_gpu_wgrad_tc.py -q > gpurun_out/t24_tests.log 2>&1; tail -2 gpurun_out/t24_tests.log)
