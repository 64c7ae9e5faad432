# round 2, job u: per-level RPN NMS, cluster-parallel top-k, residual prefetch in the row epilogue - tests, bench, busy list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_detector.py tests/test_gpu_boundary.py tests/test_gpu_conv_tc.py tests/test_gpu_ttt_step.py tests/test_gpu_entry.py -q --tb=short > gpurun_out/r02u_tests.log 2>&1; tail -3 gpurun_out/r02u_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02u_tests.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02u_bench.json 2>gpurun_out/r02u_bench.err; cut -c1-200 gpurun_out/r02u_bench.json; tail -3 gpurun_out/r02u_bench.err
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02u_busy.csv 2>gpurun_out/r02u_busy_err.log; head -2 gpurun_out/r02u_busy.csv | cut -c1-160; grep -E "rpn_topk|nms_|sort_cand|top_cand|roi_align" gpurun_out/r02u_busy.csv | cut -c1-160
TTDG_SEL_CLUSTER=1 timeout 300 python tools/run_kernels.py busy 3 gaps 2>/dev/null | grep -E "rpn_topk" | cut -c1-160
timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02u_layers_fp32.csv 2>/dev/null; head -1 gpurun_out/r02u_layers_fp32.csv; grep -E "1 1 0 8 128 128 64 256|1 1 0 8 64 64 128 512|1 1 0 8 32 32 256 1024" gpurun_out/r02u_layers_fp32.csv | cut -c1-120
