# round 2, job a: parity on BASELINE's shapes + full-step smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_configs.py -q --tb=short -s > gpurun_out/r02a_parity.log 2>&1; tail -5 gpurun_out/r02a_parity.log; grep -E "^(FAILED|E  )" gpurun_out/r02a_parity.log | cut -c1-400 | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; tail -3 gpurun_out/r02a_smoke.log
