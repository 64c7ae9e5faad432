# round 2, job ga11: dense products: two inlined row-count variants, 8 columns per batch - solver tests + fixed inputs
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02ga11_solver.log 2>&1; tail -2 gpurun_out/r02ga11_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga11_solver.log | cut -c1-300 | head
export TTDG_FIXED_MODE3=1
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-900 > gpurun_out/r02ga11_fixed.txt; cat gpurun_out/r02ga11_fixed.txt
