# round 2, job ga10: dense products out of line + 8 columns per batch: fixed inputs first, then the suite + bench of ga9
mkdir -p gpurun_out
export TTDG_FIXED_MODE3=1
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-900 > gpurun_out/r02ga10_fixed.txt; cat gpurun_out/r02ga10_fixed.txt
unset TTDG_FIXED_MODE3
bash tools/gpu_job_r02_ga9.sh
