# round 2, job y: TMA epilogue (TMA store + TMA residual load) - bit identity, timelines, per-geometry timings, layer table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q --tb=short -x > gpurun_out/r02y_tests.log 2>&1; tail -3 gpurun_out/r02y_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02y_tests.log | cut -c1-300 | head -20
(TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 4
TTDG_TRACE=1 timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 4) > gpurun_out/r02y_trace.txt 2>&1
for epi in 2 3; do
echo "== epi $epi"
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 64 256 1 0 1 0 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 64 1 0 1 0 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 64 64 128 512 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 32 32 256 1024 1 0 1 1 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 8 128 128 256 256 3 1 1 0 1 6 | cut -c60-
TTDG_TC_EPI=$epi timeout 120 python tools/conv_layer.py 800 14 14 256 1024 1 0 1 0 1 6 | cut -c60-
done
for epi in 2 3; do TTDG_TC_EPI=$epi timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02y_layers_fp32_epi$epi.csv 2>/dev/null; head -1 gpurun_out/r02y_layers_fp32_epi$epi.csv; done
