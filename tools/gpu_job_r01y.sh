mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgm_solver.py tests/test_gpu_mgm_ops.py tests/test_gpu_ttt_step.py tests/test_gpu_detector.py -m gpu -q --tb=short -x > gpurun_out/test_mgm.log 2>&1; tail -3 gpurun_out/test_mgm.log; grep -E "^(FAILED|E  )" gpurun_out/test_mgm.log | cut -c1-250 | head -20
timeout 300 python tools/run_kernels.py layers 4 12 > gpurun_out/layers_y.csv 2>gpurun_out/layers_y_err.log; head -14 gpurun_out/layers_y.csv | cut -c1-150
timeout 300 python tools/run_kernels.py timing 5 2>&1 | tail -1
