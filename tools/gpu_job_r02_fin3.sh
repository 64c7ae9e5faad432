# round 2, final gate: whole GPU suite + smoke on the closing commit
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02fin3_test_all.log 2>&1; tail -3 gpurun_out/r02fin3_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02fin3_test_all.log | cut -c1-300 | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fin3_smoke.log 2>&1; tail -2 gpurun_out/r02fin3_smoke.log | cut -c1-300
