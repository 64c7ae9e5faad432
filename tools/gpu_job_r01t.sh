mkdir -p gpurun_out
echo "== normal"; TTDG_TC_CLUSTER=1 timeout 200 python tools/run_kernels.py conv 5 2>&1 | grep conv
echo "== skip Blo"; TTDG_DEBUG_SKIP_BLO=1 TTDG_TC_CLUSTER=1 timeout 200 python tools/run_kernels.py conv 5 2>&1 | grep conv
echo "== tf32 single"; TTDG_CONV=tf32 TTDG_TC_CLUSTER=1 timeout 200 python tools/run_kernels.py conv 5 2>&1 | grep conv
echo "== cluster 2"; TTDG_TC_CLUSTER=2 timeout 200 python tools/run_kernels.py conv 5 2>&1 | grep conv
