mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -3 gpurun_out/test_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full_r01.json 2> gpurun_out/bench_err.log; tail -c 2500 gpurun_out/bench_full_r01.json; tail -5 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 3000 --csv --log-file gpurun_out/launches_full_step.csv python tools/run_kernels.py full_step 2 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 150 -c 3 -o gpurun_out/prof_conv -f python tools/run_kernels.py full_step 1 > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
ls -la gpurun_out | tail -8
