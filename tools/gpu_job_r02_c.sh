# round 2, job c: GA-GM cycle accounting + the restructured bench.py (both arms, configs 1 and 3)
mkdir -p gpurun_out
timeout 300 python tools/run_kernels.py gagm_fixed 3 > gpurun_out/r02c_gagm_fixed.log 2>&1; cat gpurun_out/r02c_gagm_fixed.log | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2>gpurun_out/r02c_bench.err; cut -c1-300 gpurun_out/r02c_bench.json; tail -3 gpurun_out/r02c_bench.err
timeout 900 python bench.py --steps 5 --warmup 3 --config 3 > gpurun_out/r02c_bench_cfg3.json 2>gpurun_out/r02c_bench_cfg3.err; cut -c1-300 gpurun_out/r02c_bench_cfg3.json; tail -3 gpurun_out/r02c_bench_cfg3.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_ref.json 2>gpurun_out/r02c_bench_ref.err; cut -c1-300 gpurun_out/r02c_bench_ref.json; tail -3 gpurun_out/r02c_bench_ref.err
