# round 2, job t: per-layer tables of a real step under the three epilogue policies, fp32 and bf16
mkdir -p gpurun_out
for epi in 0 1 2; do
TTDG_TC_EPI=$epi timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02t_layers_fp32_epi$epi.csv 2>/dev/null; head -1 gpurun_out/r02t_layers_fp32_epi$epi.csv
TTDG_CONV=bf16 TTDG_TC_EPI=$epi timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02t_layers_bf16_epi$epi.csv 2>/dev/null; head -1 gpurun_out/r02t_layers_bf16_epi$epi.csv
done
