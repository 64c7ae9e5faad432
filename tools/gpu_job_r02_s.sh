# round 2, job s: straight-line phase 2 - bit-identity tests, timelines, per-geometry timings in both epilogue modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bf16.py -q --tb=short > gpurun_out/r02s_tests.log 2>&1; tail -3 gpurun_out/r02s_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02s_tests.log | cut -c1-300 | head -20
(TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 4
TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 4
TTDG_TRACE=1 TTDG_CONV=bf16 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 4) > gpurun_out/r02s_trace.txt 2>&1
for epi in 0 1; do
echo "== epi $epi"
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 64 1 0 1 0 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 64 64 128 512 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 32 32 256 1024 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 64 64 128 512 1 0 1 1 1 6 | cut -c60-
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 6 | cut -c60-
done
