# round 2, job ga13: all dense products of the Sinkhorn-stage iterations (X = A U, V1 = A Q, V2 = W U) on the FP64 tensor cores - solver + MGM + step tests, fixed inputs, bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mgm_solver.py tests/test_gpu_mgm_ops.py tests/test_gpu_ttt_step.py -q --tb=short -x --timeout 120 > gpurun_out/r02ga13_tests.log 2>&1; tail -2 gpurun_out/r02ga13_tests.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga13_tests.log | cut -c1-300 | head
export TTDG_FIXED_MODE3=1
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3 sizes" | cut -c1-60,150-420 > gpurun_out/r02ga13_fixed_dmma.txt; cat gpurun_out/r02ga13_fixed_dmma.txt
unset TTDG_FIXED_MODE3
timeout 600 python bench.py > gpurun_out/r02ga13_bench.json 2>gpurun_out/r02ga13_bench.err; cut -c1-200 gpurun_out/r02ga13_bench.json; tail -2 gpurun_out/r02ga13_bench.err
