# round 2, job ov1: evaluation pass inside the solver windows (OverlappedEval): equivalence test, trainer tests, bench with / without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_overlap.py tests/test_gpu_ttt_step.py tests/test_gpu_entry.py -q --tb=short -x --timeout 300 > gpurun_out/r02ov1_tests.log 2>&1; tail -3 gpurun_out/r02ov1_tests.log | cut -c1-300; grep -E "^(FAILED|E  )" gpurun_out/r02ov1_tests.log | cut -c1-400 | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ov1_bench.json 2>gpurun_out/r02ov1_bench.err; cut -c1-300 gpurun_out/r02ov1_bench.json; tail -5 gpurun_out/r02ov1_bench.err
TTDG_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ov1_bench_seq.json 2>gpurun_out/r02ov1_bench_seq.err; cut -c1-300 gpurun_out/r02ov1_bench_seq.json; tail -5 gpurun_out/r02ov1_bench_seq.err
