mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -4 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-250 | head -30
timeout 300 python tools/run_kernels.py timing 5 2>&1 | tail -1
timeout 300 python tools/run_kernels.py busy 3 > gpurun_out/busy_full_step.csv 2>gpurun_out/busy_err.log; head -40 gpurun_out/busy_full_step.csv | cut -c1-150
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_m.json 2>gpurun_out/bench_m.err; cat gpurun_out/bench_m.json | cut -c1-1500
