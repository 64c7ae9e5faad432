mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/test_persist.log 2>&1; tail -3 gpurun_out/test_persist.log; grep -E "^(FAILED|E  )" gpurun_out/test_persist.log | cut -c1-300 | head -20
timeout 600 python -m pytest tests/test_gpu_detector.py tests/test_gpu_ttt_step.py -m gpu -q --tb=short > gpurun_out/test_persist2.log 2>&1; tail -2 gpurun_out/test_persist2.log; grep -E "^(FAILED|E  )" gpurun_out/test_persist2.log | cut -c1-300 | head -20
timeout 200 python tools/run_kernels.py conv 5 8,128,128,256,256,3 8,32,32,256,256,3 800,14,14,256,256,3 8000,1,1,12544,1024,1 8,128,128,64,256,1 8,64,64,128,512,1 8,32,32,256,1024,1 8,128,128,256,64,1 8,128,128,256,256,1 800,28,28,256,64,1 8,16,16,512,512,3 2>&1 | grep conv
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ad.json 2>gpurun_out/bench_ad.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ad.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['gpu_launches']);print(d['roofline_step_dominant'])"
