# round 2, job m: run-length radix select, 16-byte bf16 epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_bf16.py tests/test_gpu_conv_tc.py -q --tb=short > gpurun_out/r02m_tests.log 2>&1; tail -3 gpurun_out/r02m_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02m_tests.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02m_bench.json 2>gpurun_out/r02m_bench.err; cut -c1-200 gpurun_out/r02m_bench.json; tail -3 gpurun_out/r02m_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --config 2 > gpurun_out/r02m_bench_cfg2.json 2>gpurun_out/r02m_bench_cfg2.err; cut -c1-200 gpurun_out/r02m_bench_cfg2.json
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02m_busy.csv 2>gpurun_out/r02m_busy_err.log; head -2 gpurun_out/r02m_busy.csv | cut -c1-160; grep -E "rpn_topk|nms_sweep|sort_cand" gpurun_out/r02m_busy.csv | head -4 | cut -c1-160
TTDG_CONV=bf16 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02m_layers_bf16.csv 2>/dev/null; head -14 gpurun_out/r02m_layers_bf16.csv | cut -c1-150
