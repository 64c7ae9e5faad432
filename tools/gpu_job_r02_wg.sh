# round 2, job wg: persistent weight-gradient kernel with the TMA reduce-add epilogue - tests, per-layer table, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_detector.py tests/test_gpu_bf16.py tests/test_gpu_parity_configs.py -q --tb=short -x > gpurun_out/r02wg_tests.log 2>&1; tail -3 gpurun_out/r02wg_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02wg_tests.log | cut -c1-300 | head -20
timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02wg_layers_fp32.csv 2>/dev/null; head -1 gpurun_out/r02wg_layers_fp32.csv; grep wgrad gpurun_out/r02wg_layers_fp32.csv | head -12 | cut -c1-110
TTDG_CONV=bf16 timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02wg_layers_bf16.csv 2>/dev/null; head -1 gpurun_out/r02wg_layers_bf16.csv
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02wg_bench.json 2>gpurun_out/r02wg_bench.err; cut -c1-200 gpurun_out/r02wg_bench.json; tail -3 gpurun_out/r02wg_bench.err
