# round 2, closing job a: whole GPU suite, smoke, the three bench configs + the reference arm, live kernel shares, stage timing
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02fa_test_all.log 2>&1; tail -3 gpurun_out/r02fa_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02fa_test_all.log | cut -c1-300 | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fa_smoke.log 2>&1; tail -2 gpurun_out/r02fa_smoke.log | cut -c1-300
for c in 1 2 3; do timeout 900 python bench.py --steps 20 --warmup 3 --config $c > gpurun_out/r02fa_bench_cfg$c.json 2>gpurun_out/r02fa_bench_cfg$c.err; cut -c1-200 gpurun_out/r02fa_bench_cfg$c.json; done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02fa_bench_reference.json 2>gpurun_out/r02fa_bench_reference.err; cut -c1-200 gpurun_out/r02fa_bench_reference.json
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/r02fa_busy.csv 2>/dev/null; head -2 gpurun_out/r02fa_busy.csv | cut -c1-160
timeout 300 python tools/run_kernels.py timing 3 > gpurun_out/r02fa_timing.log 2>&1; tail -1 gpurun_out/r02fa_timing.log | cut -c1-500
timeout 300 python tools/run_kernels.py layers 3 90 > gpurun_out/r02fa_layers_fp32.csv 2>/dev/null; head -1 gpurun_out/r02fa_layers_fp32.csv
