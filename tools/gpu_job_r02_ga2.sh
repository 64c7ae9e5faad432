# round 2, job ga2: GA-GM Hungarian-stage iterations out of shared memory (hfast) - solver tests, then A/B on the bench workload
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02ga2_solver.log 2>&1; tail -4 gpurun_out/r02ga2_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga2_solver.log | cut -c1-300 | head
TTDG_GAGM_HFAST=1 timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-600 > gpurun_out/r02ga2_hfast1.txt
TTDG_GAGM_HFAST=0 timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-600 > gpurun_out/r02ga2_hfast0.txt
cat gpurun_out/r02ga2_hfast1.txt gpurun_out/r02ga2_hfast0.txt
