# round 2, closing check after the last host-side changes: overlap test + the default bench command
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_overlap.py -q --tb=short -x --timeout 200 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02fin2_bench_default.json 2>gpurun_out/r02fin2_bench_default.err; cut -c1-260 gpurun_out/r02fin2_bench_default.json; tail -2 gpurun_out/r02fin2_bench_default.err
