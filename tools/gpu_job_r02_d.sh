# round 2, job d: whole GPU suite after the boundary refactor + bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02d_test_all.log 2>&1; tail -4 gpurun_out/r02d_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02d_test_all.log | cut -c1-300 | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2>gpurun_out/r02d_bench.err; cut -c1-200 gpurun_out/r02d_bench.json; tail -3 gpurun_out/r02d_bench.err
