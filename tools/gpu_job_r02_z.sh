# round 2, job z: TMA epilogue with bf16 tiles - bit identity, bf16 parity, timings, per-layer tables, bench configs 1 / 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bf16.py tests/test_gpu_detector.py -q --tb=short -x > gpurun_out/r02z_tests.log 2>&1; tail -3 gpurun_out/r02z_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02z_tests.log | cut -c1-300 | head -20
for epi in 2 3; do
echo "== bf16 epi $epi"
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 64 64 128 512 1 0 1 1 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 32 32 256 1024 1 0 1 1 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 6 | cut -c60-
done
for epi in 2 3; do TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02z_layers_bf16_epi$epi.csv 2>/dev/null; head -1 gpurun_out/r02z_layers_bf16_epi$epi.csv; done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02z_bench.json 2>gpurun_out/r02z_bench.err; cut -c1-200 gpurun_out/r02z_bench.json; tail -3 gpurun_out/r02z_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --config 2 > gpurun_out/r02z_bench_cfg2.json 2>gpurun_out/r02z_bench_cfg2.err; cut -c1-200 gpurun_out/r02z_bench_cfg2.json
