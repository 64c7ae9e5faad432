# round 2, job v: whole GPU suite (device resize, per-level NMS, cluster top-k, epilogues, CfgNode) + smoke + bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02v_test_all.log 2>&1; tail -3 gpurun_out/r02v_test_all.log; grep -E "^(FAILED|E  )" gpurun_out/r02v_test_all.log | cut -c1-300 | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; tail -2 gpurun_out/r02v_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02v_bench.json 2>gpurun_out/r02v_bench.err; cut -c1-200 gpurun_out/r02v_bench.json; tail -3 gpurun_out/r02v_bench.err
