mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -3 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-300 | head -20
timeout 500 python tools/run_kernels.py busy 3 gaps 2>/dev/null | grep -A8 "^wall" | cut -c1-200
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ab.json 2>gpurun_out/bench_ab.err; tail -3 gpurun_out/bench_ab.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ab.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['gpu_launches']);print(d['roofline_step_dominant'])"
