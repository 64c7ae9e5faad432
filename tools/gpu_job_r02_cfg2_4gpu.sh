# round 2 (gpurun --gpus 4): BASELINE.json configs[2] as named - 32 images over 4 B200s, bf16 backbone, fp32 / fp64 matching stage
mkdir -p gpurun_out
N=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --config 2 > gpurun_out/r02_bench_cfg2_${N}gpu.json 2>gpurun_out/r02_bench_cfg2_${N}gpu.err; cut -c1-300 gpurun_out/r02_bench_cfg2_${N}gpu.json; tail -2 gpurun_out/r02_bench_cfg2_${N}gpu.err | cut -c1-200
