mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_ttt_step.py tests/test_gpu_entry.py -m gpu -q --tb=short > gpurun_out/test_nms.log 2>&1; tail -3 gpurun_out/test_nms.log; grep -E "^(FAILED|E  )" gpurun_out/test_nms.log | cut -c1-300 | head -20
timeout 300 python tools/run_kernels.py layers 3 70 2>/dev/null | grep "ttdg_nms\|sum of" | cut -c1-150
TTDG_NMS_FUSED=0 timeout 300 python tools/run_kernels.py layers 3 70 2>/dev/null | grep "ttdg_nms\|sum of" | cut -c1-150
