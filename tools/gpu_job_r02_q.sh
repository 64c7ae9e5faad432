# round 2, job q: + first drain chunk straight into the accumulators, FrozenBN scale / bias in phase 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bf16.py -q --tb=short > gpurun_out/r02q_tests.log 2>&1; tail -3 gpurun_out/r02q_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02q_tests.log | cut -c1-300 | head -20
for epi in 0 1; do
echo "== epi $epi"
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 64 1 0 1 0 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 64 64 128 512 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 6 | cut -c60-
done
TTDG_TC_EPI=1 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02q_layers_fp32_epi1.csv 2>/dev/null; head -1 gpurun_out/r02q_layers_fp32_epi1.csv
TTDG_CONV=bf16 TTDG_TC_EPI=1 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02q_layers_bf16_epi1.csv 2>/dev/null; head -1 gpurun_out/r02q_layers_bf16_epi1.csv
TTDG_CONV=bf16 TTDG_TC_EPI=0 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02q_layers_bf16_epi0.csv 2>/dev/null; head -1 gpurun_out/r02q_layers_bf16_epi0.csv
