# round 2, job h: label-correcting LAP augmentations; selection test fix; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgm_solver.py tests/test_gpu_ttt_step.py tests/test_gpu_detector.py -q --tb=short > gpurun_out/r02h_tests.log 2>&1; tail -4 gpurun_out/r02h_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02h_tests.log | cut -c1-300 | head -30
timeout 300 python tools/run_kernels.py gagm_fixed 3 > gpurun_out/r02h_gagm_fixed.log 2>&1; grep "lap_fast [03]" gpurun_out/r02h_gagm_fixed.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02h_bench.json 2>gpurun_out/r02h_bench.err; cut -c1-250 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02h_bench2.json 2>gpurun_out/r02h_bench2.err; cut -c1-250 gpurun_out/r02h_bench2.json
