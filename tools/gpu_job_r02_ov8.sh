# round 2, job ov8 (gpurun --gpus 8): the overlapped schedule on 8 ranks, configs[1]
mkdir -p gpurun_out
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02ov8_bench_${N}gpu.json 2>gpurun_out/r02ov8_bench_${N}gpu.err; cut -c1-300 gpurun_out/r02ov8_bench_${N}gpu.json; tail -3 gpurun_out/r02ov8_bench_${N}gpu.err | cut -c1-300
