# round 2, job f: LAP two-redux step, bf16 heads, fixed tests, microbench with affinity N = 256..1024
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgm_solver.py tests/test_gpu_bf16.py tests/test_gpu_boundary.py tests/test_gpu_entry.py tests/test_gpu_ttt_step.py tests/test_gpu_mgm_ops.py -q --tb=short > gpurun_out/r02f_tests.log 2>&1; tail -4 gpurun_out/r02f_tests.log; grep -E "^(FAILED|E  )" gpurun_out/r02f_tests.log | cut -c1-300 | head -30
timeout 300 python tools/run_kernels.py gagm_fixed 3 > gpurun_out/r02f_gagm_fixed.log 2>&1; grep "lap_fast 3" gpurun_out/r02f_gagm_fixed.log | cut -c1-400
timeout 600 python tools/sinkhorn_microbench.py > gpurun_out/r02f_microbench.jsonl 2>gpurun_out/r02f_microbench.err; cat gpurun_out/r02f_microbench.jsonl | cut -c1-330; tail -2 gpurun_out/r02f_microbench.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 2 > gpurun_out/r02f_bench_cfg2.json 2>gpurun_out/r02f_bench_cfg2.err; cut -c1-250 gpurun_out/r02f_bench_cfg2.json; tail -3 gpurun_out/r02f_bench_cfg2.err
timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02f_layers_fp32.csv 2>gpurun_out/r02f_layers_err.log; head -3 gpurun_out/r02f_layers_fp32.csv | cut -c1-150
TTDG_CONV=bf16 timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/r02f_layers_bf16.csv 2>gpurun_out/r02f_layers_bf16_err.log; head -40 gpurun_out/r02f_layers_bf16.csv | cut -c1-150
