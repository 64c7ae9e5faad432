# round 2, job w: GA-GM Hungarian-stage iteration with batched gathers / norm loads - solver tests, fixed-input timings, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgm_solver.py tests/test_gpu_ttt_step.py tests/test_gpu_entry.py tests/test_resize.py -m gpu -q --tb=short > gpurun_out/r02w_tests.log 2>&1; tail -3 gpurun_out/r02w_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02w_tests.log | cut -c1-300 | head
timeout 300 python tools/run_kernels.py gagm_fixed 3 > gpurun_out/r02w_gagm_fixed.log 2>&1; cut -c1-330 gpurun_out/r02w_gagm_fixed.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02w_bench.json 2>gpurun_out/r02w_bench.err; cut -c1-200 gpurun_out/r02w_bench.json; tail -3 gpurun_out/r02w_bench.err
