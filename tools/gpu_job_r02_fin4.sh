# round 2, last check: the default bench command on the closing commit
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r02fin4_bench_default.json 2>gpurun_out/r02fin4_bench_default.err; cut -c1-200 gpurun_out/r02fin4_bench_default.json; tail -2 gpurun_out/r02fin4_bench_default.err
