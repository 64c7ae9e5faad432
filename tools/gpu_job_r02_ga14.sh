# round 2, job ga14: 32 k-steps per load batch in the W U tensor-core product - solver tests + fixed inputs + the solver on the bench workload
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02ga14_solver.log 2>&1; tail -2 gpurun_out/r02ga14_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga14_solver.log | cut -c1-300 | head
export TTDG_FIXED_MODE3=1
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3 sizes" | cut -c1-60,150-420 > gpurun_out/r02ga14_fixed.txt; cat gpurun_out/r02ga14_fixed.txt
unset TTDG_FIXED_MODE3
timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep "gagm_bench step" | cut -c1-420 > gpurun_out/r02ga14_gagm_bench.txt; cat gpurun_out/r02ga14_gagm_bench.txt
