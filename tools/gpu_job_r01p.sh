mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/test_conv.log 2>&1; tail -4 gpurun_out/test_conv.log; grep -E "^(FAILED|E  )" gpurun_out/test_conv.log | cut -c1-250 | head -20
timeout 300 python tools/run_kernels.py layers 3 30 > gpurun_out/layers_ts.csv 2>gpurun_out/layers_ts_err.log; head -32 gpurun_out/layers_ts.csv | cut -c1-160
