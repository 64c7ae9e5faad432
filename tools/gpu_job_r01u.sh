mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/test_conv_2g.log 2>&1; tail -2 gpurun_out/test_conv_2g.log; grep -E "^(FAILED|E  )" gpurun_out/test_conv_2g.log | cut -c1-250 | head -10
echo "== cluster 1"; TTDG_TC_CLUSTER=1 timeout 200 python tools/run_kernels.py conv 5 2>&1 | grep conv
echo "== cluster 2"; TTDG_TC_CLUSTER=2 timeout 200 python tools/run_kernels.py conv 5 2>&1 | grep conv
echo "== cluster 4"; TTDG_TC_CLUSTER=4 timeout 200 python tools/run_kernels.py conv 5 2>&1 | grep conv
TTDG_TC_CLUSTER=1 timeout 300 python tools/run_kernels.py layers 3 40 > gpurun_out/layers_2g.csv 2>gpurun_out/layers_2g_err.log; grep "wgrad_tc\|sum of" gpurun_out/layers_2g.csv | head -8 | cut -c1-160
