mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test.log 2>&1; tail -3 gpurun_out/test.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_err.log; tail -c 3000 gpurun_out/bench_r01.json; tail -5 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01.json 2>&1; tail -c 600 gpurun_out/bench_ref_r01.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mgm_step.csv python tools/run_kernels.py mgm_step 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_stream -s 1 -c 1 -o gpurun_out/prof_sinkhorn_stream -f python tools/run_kernels.py sinkhorn_stream 2 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gagm -s 1 -c 1 -o gpurun_out/prof_gagm -f python tools/run_kernels.py gagm 2 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ls -la gpurun_out
