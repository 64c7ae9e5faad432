mkdir -p gpurun_out
TTDG_TC_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/test_conv_wg.log 2>&1; tail -2 gpurun_out/test_conv_wg.log; grep -E "^(FAILED|E  )" gpurun_out/test_conv_wg.log | cut -c1-250 | head -10
echo "== normal (cluster 1)"
TTDG_TC_CLUSTER=1 timeout 300 python tools/run_kernels.py layers 3 40 > gpurun_out/layers_wgts.csv 2>gpurun_out/layers_wgts_err.log; grep "wgrad_tc\|sum of" gpurun_out/layers_wgts.csv | head -12 | cut -c1-160
echo "== skip Blo (cluster 1)"
TTDG_DEBUG_SKIP_BLO=1 TTDG_TC_CLUSTER=1 timeout 300 python tools/run_kernels.py layers 3 26 > gpurun_out/layers_skipblo.csv 2>gpurun_out/layers_skipblo_err.log; grep "conv_tc\|sum of" gpurun_out/layers_skipblo.csv | head -10 | cut -c1-160
