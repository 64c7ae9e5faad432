mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2>gpurun_out/bench_2gpu.err; echo "rc=$?"; wc -c gpurun_out/bench_2gpu.json; grep -v "^\*\|OMP_NUM" gpurun_out/bench_2gpu.err | tail -25 | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/bench_2gpu.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['roofline_step_dominant']['achieved'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>gpurun_out/ref2.err | cut -c1-200; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/ref2.err | tail -5 | cut -c1-300
