# round 2, job j: LAP modes on the bench workload's problems, weight refresh, full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short > gpurun_out/r02j_solver.log 2>&1; tail -3 gpurun_out/r02j_solver.log; grep -E "^(FAILED|E  )" gpurun_out/r02j_solver.log | cut -c1-300 | head -10
timeout 600 python tools/run_kernels.py gagm_bench 3 > gpurun_out/r02j_gagm_bench.log 2>&1; grep "gagm_bench" gpurun_out/r02j_gagm_bench.log | cut -c1-330; tail -2 gpurun_out/r02j_gagm_bench.log | cut -c1-300
timeout 1800 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_mgm_solver.py > gpurun_out/r02j_test_all.log 2>&1; tail -3 gpurun_out/r02j_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02j_test_all.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02j_bench.json 2>gpurun_out/r02j_bench.err; cut -c1-250 gpurun_out/r02j_bench.json; tail -3 gpurun_out/r02j_bench.err
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02j_busy.csv 2>gpurun_out/r02j_busy_err.log; head -4 gpurun_out/r02j_busy.csv | cut -c1-160
