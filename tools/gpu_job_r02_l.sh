# round 2, job l: faster radix select, fused ReLU/BN backward - whole suite + bench + launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02l_test_all.log 2>&1; tail -3 gpurun_out/r02l_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02l_test_all.log | cut -c1-300 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1; tail -2 gpurun_out/r02l_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02l_bench.json 2>gpurun_out/r02l_bench.err; cut -c1-200 gpurun_out/r02l_bench.json; tail -3 gpurun_out/r02l_bench.err
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02l_busy.csv 2>gpurun_out/r02l_busy_err.log; head -3 gpurun_out/r02l_busy.csv | cut -c1-160; grep -E "rpn_topk|relu_bn|nms_sweep|sort_cand" gpurun_out/r02l_busy.csv | head -8 | cut -c1-160
