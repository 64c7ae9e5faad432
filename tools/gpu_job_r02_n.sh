# round 2, job n: warp-transposed (coalesced) conv epilogue - bit-identity test, per-layer A/B tables, bench configs 1/2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bf16.py -q --tb=short > gpurun_out/r02n_tests.log 2>&1; tail -3 gpurun_out/r02n_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02n_tests.log | cut -c1-300 | head -20
TTDG_TC_EPI=0 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02n_layers_fp32_epi0.csv 2>/dev/null; head -1 gpurun_out/r02n_layers_fp32_epi0.csv
TTDG_TC_EPI=1 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02n_layers_fp32_epi1.csv 2>/dev/null; head -1 gpurun_out/r02n_layers_fp32_epi1.csv
TTDG_CONV=bf16 TTDG_TC_EPI=0 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02n_layers_bf16_epi0.csv 2>/dev/null; head -1 gpurun_out/r02n_layers_bf16_epi0.csv
TTDG_CONV=bf16 TTDG_TC_EPI=1 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02n_layers_bf16_epi1.csv 2>/dev/null; head -1 gpurun_out/r02n_layers_bf16_epi1.csv
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02n_bench.json 2>gpurun_out/r02n_bench.err; cut -c1-200 gpurun_out/r02n_bench.json; tail -3 gpurun_out/r02n_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --config 2 > gpurun_out/r02n_bench_cfg2.json 2>gpurun_out/r02n_bench_cfg2.err; cut -c1-200 gpurun_out/r02n_bench_cfg2.json
