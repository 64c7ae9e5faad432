# round 2, job ga3: GA-GM without F2F in the products (integer widening, fp64 copy of the own block of A), DSMEM exchange, LAP phase cycles
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02ga3_solver.log 2>&1; tail -4 gpurun_out/r02ga3_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga3_solver.log | cut -c1-300 | head
timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-700 > gpurun_out/r02ga3_default.txt
TTDG_GAGM_FASTCVT=0 timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-700 > gpurun_out/r02ga3_f2f.txt
TTDG_GAGM_HFAST=2 timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-700 > gpurun_out/r02ga3_nocache.txt
cat gpurun_out/r02ga3_default.txt gpurun_out/r02ga3_f2f.txt gpurun_out/r02ga3_nocache.txt
