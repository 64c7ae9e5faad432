# round 2, job ga8: F2F widening restored; lean LAP inlined (A) vs __noinline__ (B, -DTTDG_LAP_NOINLINE) on the fixed inputs
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02ga8_solver.log 2>&1; tail -2 gpurun_out/r02ga8_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga8_solver.log | cut -c1-300 | head
export TTDG_FIXED_MODE3=1
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-900 > gpurun_out/r02ga8_a.txt
TTDG_LIB=$PWD/ttdg-mgm_b200/ttdg_b200/lib/libttdg_sm100_b.so timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-900 > gpurun_out/r02ga8_b.txt
for f in a b; do echo "== $f"; cat gpurun_out/r02ga8_$f.txt; done
