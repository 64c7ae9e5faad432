mkdir -p gpurun_out
TTDG_CONV=tf32 timeout 300 python tools/run_kernels.py layers 3 40 > gpurun_out/layers_tf32.csv 2>gpurun_out/layers_tf32_err.log; head -42 gpurun_out/layers_tf32.csv | cut -c1-160
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_o.json 2>gpurun_out/bench_o.err; cut -c1-300 gpurun_out/bench_o.json; python -c "
import json;d=json.load(open('gpurun_out/bench_o.json'));print(d['value'],d['e2e'],d['gagm'])"
