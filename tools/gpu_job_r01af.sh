mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -3 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-300 | head -20
timeout 200 python tools/run_kernels.py conv 5 800,14,14,256,256,3 800,28,28,256,64,1 800,14,14,256,1024,1 2>&1 | grep conv
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['gpu_launches']);print(d['roofline_step_dominant']['achieved'], d['roofline_step_dominant']['ms_per_step'])"
timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/layers_final.csv 2>gpurun_out/layers_final_err.log; head -3 gpurun_out/layers_final.csv | cut -c1-150
timeout 300 python tools/run_kernels.py busy 3 > gpurun_out/busy_final.csv 2>gpurun_out/busy_final_err.log; head -2 gpurun_out/busy_final.csv | cut -c1-150
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -s 1100 -c 941 --csv --log-file gpurun_out/launches_final.csv python tools/run_kernels.py full_step 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
