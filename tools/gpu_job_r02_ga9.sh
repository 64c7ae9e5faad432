# round 2, job ga9: GA-GM state after the shared-memory Hungarian path + lean-LAP scan rewrites: whole GPU suite, bench line, solver on the bench workload
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02ga9_test_all.log 2>&1; tail -3 gpurun_out/r02ga9_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02ga9_test_all.log | cut -c1-300 | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ga9_bench.json 2>gpurun_out/r02ga9_bench.err; cut -c1-330 gpurun_out/r02ga9_bench.json; tail -3 gpurun_out/r02ga9_bench.err
timeout 200 python tools/run_kernels.py gagm_bench 2 2>&1 | grep gagm_bench | cut -c1-800 > gpurun_out/r02ga9_gagm_bench.txt; cat gpurun_out/r02ga9_gagm_bench.txt
