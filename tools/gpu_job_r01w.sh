mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_detector.py tests/test_gpu_evaluator.py -m gpu -q --tb=short > gpurun_out/test_epi.log 2>&1; tail -3 gpurun_out/test_epi.log; grep -E "^(FAILED|E  )" gpurun_out/test_epi.log | cut -c1-250 | head -30
timeout 300 python tools/run_kernels.py layers 3 60 > gpurun_out/layers_epi.csv 2>gpurun_out/layers_epi_err.log; head -50 gpurun_out/layers_epi.csv | cut -c1-150
timeout 300 python tools/run_kernels.py timing 5 2>&1 | tail -1
