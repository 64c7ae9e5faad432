# round 2, job sk2: GA-GM Sinkhorn projector in the scaling form after two log-domain steps; NaN guard.  Tight timeouts.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02sk2_solver.log 2>&1; tail -4 gpurun_out/r02sk2_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02sk2_solver.log | cut -c1-300 | head
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3" | cut -c1-60,150-420
timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-420
