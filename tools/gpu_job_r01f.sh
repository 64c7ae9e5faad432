mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -4 gpurun_out/test_all.log; grep -E "^(FAILED)" gpurun_out/test_all.log | cut -c1-200 | head -20
timeout 300 python tools/run_kernels.py timing 3 2>&1 | tail -1
