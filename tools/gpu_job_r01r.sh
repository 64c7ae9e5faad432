mkdir -p gpurun_out
for cl in 2 4; do
TTDG_TC_CLUSTER=$cl timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/test_conv_cl$cl.log 2>&1; tail -2 gpurun_out/test_conv_cl$cl.log; grep -E "^(FAILED|E  )" gpurun_out/test_conv_cl$cl.log | cut -c1-250 | head -10
done
for cl in 1 2 4; do
echo "== cluster $cl"
TTDG_TC_CLUSTER=$cl timeout 300 python tools/run_kernels.py layers 3 26 > gpurun_out/layers_cl$cl.csv 2>gpurun_out/layers_cl${cl}_err.log; grep "conv_tc\|sum of" gpurun_out/layers_cl$cl.csv | head -16 | cut -c1-160
done
