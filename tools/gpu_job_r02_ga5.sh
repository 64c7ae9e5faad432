# round 2, job ga6: lean LAP - compare+select scans, column checks of the certificate by the column lane, free rows per round
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02ga6_solver.log 2>&1; tail -4 gpurun_out/r02ga6_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga6_solver.log | cut -c1-300 | head
timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-700 > gpurun_out/r02ga6_default.txt
cat gpurun_out/r02ga6_default.txt
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-420
