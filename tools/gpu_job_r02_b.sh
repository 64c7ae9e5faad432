# round 2, job b: lean certified LAP (mode 3) - solver tests, timings, parity, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mgm_solver.py tests/test_gpu_ttt_step.py -q --tb=short -x > gpurun_out/r02b_solver.log 2>&1; tail -5 gpurun_out/r02b_solver.log; grep -E "^(FAILED|E  )" gpurun_out/r02b_solver.log | cut -c1-300 | head -20
timeout 300 python tools/run_kernels.py gagm_fixed 5 > gpurun_out/r02b_gagm_fixed.log 2>&1; cat gpurun_out/r02b_gagm_fixed.log | cut -c1-260
timeout 1500 python -m pytest tests/test_gpu_parity_configs.py -q --tb=short -s > gpurun_out/r02b_parity.log 2>&1; tail -3 gpurun_out/r02b_parity.log; grep -E "^(FAILED|E  )" gpurun_out/r02b_parity.log | cut -c1-400 | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -3 gpurun_out/r02b_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2>gpurun_out/r02b_bench.err; cut -c1-400 gpurun_out/r02b_bench.json
TTDG_LAP_FAST=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_lap0.json 2>gpurun_out/r02b_bench_lap0.err; cut -c1-300 gpurun_out/r02b_bench_lap0.json
