# round 2, job ga7: deterministic A/B on the fixed inputs: default / without the shared-memory Hungarian path / without integer widening
mkdir -p gpurun_out
export TTDG_FIXED_MODE3=1
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-700 > gpurun_out/r02ga7_default.txt
TTDG_GAGM_HFAST=0 timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-700 > gpurun_out/r02ga7_hfast0.txt
TTDG_GAGM_HFAST=0 TTDG_GAGM_FASTCVT=0 timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-700 > gpurun_out/r02ga7_hfast0_f2f.txt
for f in default hfast0 hfast0_f2f; do echo "== $f"; cat gpurun_out/r02ga7_$f.txt; done
