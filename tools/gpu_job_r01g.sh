mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/test_tc.log 2>&1; tail -15 gpurun_out/test_tc.log | cut -c1-250
timeout 300 python tools/run_kernels.py timing 3 2>&1 | tail -1
TTDG_WGRAD_TC=0 timeout 300 python tools/run_kernels.py timing 3 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -4 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-250 | head -20
