# round 2, job ga12: W U product of the Sinkhorn-stage iterations on the FP64 tensor cores (mma.sync m8n8k4 f64) - solver tests, fixed inputs A/B
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 > gpurun_out/r02ga12_solver.log 2>&1; tail -2 gpurun_out/r02ga12_solver.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ga12_solver.log | cut -c1-300 | head
export TTDG_FIXED_MODE3=1
timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3 sizes" | cut -c1-60,150-420 > gpurun_out/r02ga12_fixed_dmma.txt; cat gpurun_out/r02ga12_fixed_dmma.txt
TTDG_GAGM_UALL=1 timeout 120 python tools/run_kernels.py gagm_fixed 2 2>&1 | grep "lap_fast 3 sizes" | cut -c1-60,150-420 > gpurun_out/r02ga12_fixed_vector.txt; cat gpurun_out/r02ga12_fixed_vector.txt
