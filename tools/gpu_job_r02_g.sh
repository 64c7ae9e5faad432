# round 2, job g: device-side selection (csrc/select.cu) - tests, parity, bench, launch counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_boundary.py tests/test_gpu_ttt_step.py tests/test_gpu_entry.py -q --tb=short > gpurun_out/r02g_tests.log 2>&1; tail -4 gpurun_out/r02g_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02g_tests.log | cut -c1-300 | head -30
timeout 900 python -m pytest tests/test_gpu_parity_configs.py tests/test_gpu_bf16.py -q --tb=short > gpurun_out/r02g_parity.log 2>&1; tail -3 gpurun_out/r02g_parity.log; grep -E "^(FAILED|E  )" gpurun_out/r02g_parity.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2>gpurun_out/r02g_bench.err; cut -c1-250 gpurun_out/r02g_bench.json; tail -3 gpurun_out/r02g_bench.err
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02g_busy.csv 2>gpurun_out/r02g_busy_err.log; head -30 gpurun_out/r02g_busy.csv | cut -c1-160
