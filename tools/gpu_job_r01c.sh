mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_ttt_step.py -m gpu -q --tb=short -x > gpurun_out/test_det.log 2>&1; tail -3 gpurun_out/test_det.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full_r01.json 2> gpurun_out/bench_err.log; tail -c 3000 gpurun_out/bench_full_r01.json; tail -5 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_full_ref_r01.json 2>&1; tail -c 400 gpurun_out/bench_full_ref_r01.json
