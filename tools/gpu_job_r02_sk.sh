# round 2, job sk: Sinkhorn step with three lines per warp in flight + redux max - solver / ops tests, fixed-input GA-GM timings, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgm_solver.py tests/test_gpu_mgm_ops.py tests/test_gpu_ttt_step.py tests/test_gpu_parity_configs.py -q --tb=short > gpurun_out/r02sk_tests.log 2>&1; tail -3 gpurun_out/r02sk_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02sk_tests.log | cut -c1-300 | head
timeout 300 python tools/run_kernels.py gagm_fixed 3 2>&1 | grep "lap_fast 3" | cut -c1-330
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02sk_bench.json 2>gpurun_out/r02sk_bench.err; cut -c1-200 gpurun_out/r02sk_bench.json; tail -3 gpurun_out/r02sk_bench.err
timeout 300 python tools/run_kernels.py busy 3 gaps 2>/dev/null | grep -E "gagm_kernel|sinkhorn_small|wall ms" | cut -c1-120
