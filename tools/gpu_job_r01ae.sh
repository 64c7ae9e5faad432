mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_detector.py -m gpu -q --tb=short > gpurun_out/test_bn64.log 2>&1; tail -2 gpurun_out/test_bn64.log; grep -E "^(FAILED|E  )" gpurun_out/test_bn64.log | cut -c1-300 | head -20
timeout 200 python tools/run_kernels.py conv 5 8,16,16,512,512,3 8,16,16,2048,512,1 8,16,16,512,2048,1 8,32,32,256,256,3 8,32,32,1024,256,1 8,16,16,2048,256,1 8,8,8,256,256,3 2>&1 | grep conv
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ae.json 2>gpurun_out/bench_ae.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ae.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['gpu_launches']);print(d['roofline_step_dominant']['achieved'], d['roofline_step_dominant']['ms_per_step'])"
