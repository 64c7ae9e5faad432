# round 2, job e: bf16 backbone + boundary + whole suite + bench configs 1/2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q --tb=short -s > gpurun_out/r02e_bf16.log 2>&1; tail -4 gpurun_out/r02e_bf16.log; grep -E "^(FAILED|E  )" gpurun_out/r02e_bf16.log | cut -c1-300 | head -30
timeout 1800 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_bf16.py > gpurun_out/r02e_test_all.log 2>&1; tail -4 gpurun_out/r02e_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02e_test_all.log | cut -c1-300 | head -30
timeout 900 python bench.py --steps 10 --warmup 3 --config 2 > gpurun_out/r02e_bench_cfg2.json 2>gpurun_out/r02e_bench_cfg2.err; cut -c1-250 gpurun_out/r02e_bench_cfg2.json; tail -3 gpurun_out/r02e_bench_cfg2.err
