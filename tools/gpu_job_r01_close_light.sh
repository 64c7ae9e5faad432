# lighter closing pass (no --set full captures): tests, smoke, two bench lines, live shares, one-period launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/test_all.log 2>&1; tail -3 gpurun_out/test_all.log; grep -E "^(FAILED|E  )" gpurun_out/test_all.log | cut -c1-300 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_close$i.json 2>gpurun_out/bench_close$i.err; python -c "
import json;d=json.load(open('gpurun_out/bench_close$i.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['gpu_launches'],d['roofline_step_dominant']['achieved'])"; done
timeout 300 python tools/run_kernels.py layers 3 70 > gpurun_out/layers_final.csv 2>gpurun_out/layers_final_err.log; head -2 gpurun_out/layers_final.csv | cut -c1-150
timeout 300 python tools/run_kernels.py busy 3 gaps > gpurun_out/busy_final.csv 2>gpurun_out/busy_final_err.log; head -3 gpurun_out/busy_final.csv | cut -c1-150
N=$(grep -o "launches/step [0-9]*" gpurun_out/busy_final.csv | grep -o "[0-9]*$"); echo "launches per step: $N"
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -s 1100 -c $N --csv --log-file gpurun_out/launches_final.csv python tools/run_kernels.py full_step 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
