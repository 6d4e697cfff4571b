mkdir -p gpurun_out/final
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final/gpu_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/final/gpu_tests.txt)
tail -3 gpurun_out/final/gpu_tests.txt
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err
timeout 500 python bench.py > gpurun_out/final/bench_c3_n1.json 2> gpurun_out/final/bench_c3_n1.err
tail -2 gpurun_out/final/bench_c3_n1.err
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/final/bench_c5_n1.json 2> gpurun_out/final/bench_c5_n1.err
timeout 300 python bench.py --workload c2 --no-heads --no-cpu-baseline > gpurun_out/final/bench_c2_n1.json 2> gpurun_out/final/bench_c2_n1.err
timeout 120 python scripts/probes/op_times.py resnet50 768 > gpurun_out/final/op_times_resnet50.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-heads --no-cpu-baseline --no-stored-e2e > gpurun_out/final/launches_c3.log 2>&1
timeout 600 ncu --set full --profile-from-start off --clock-control none -o gpurun_out/final/step_c3 python scripts/probes/one_step.py c3 32 > gpurun_out/final/ncu_full.log 2>&1
ncu -i gpurun_out/final/step_c3.ncu-rep --page raw --csv > gpurun_out/final/step_c3_raw.csv 2>/dev/null
ls -la gpurun_out/final/
rm -f gpurun_out/final/step_c3.ncu-rep
