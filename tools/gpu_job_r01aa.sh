mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_ttt_step.py tests/test_gpu_entry.py tests/test_gpu_evaluator.py -m gpu -q --tb=short > gpurun_out/test_det.log 2>&1; tail -3 gpurun_out/test_det.log; grep -E "^(FAILED|E  )" gpurun_out/test_det.log | cut -c1-300 | head -20
timeout 500 python tools/run_kernels.py busy 3 gaps 2>/dev/null | grep -A22 "^wall" | cut -c1-200
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_aa.json 2>gpurun_out/bench_aa.err; python -c "
import json;d=json.load(open('gpurun_out/bench_aa.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['gpu_launches'])"
