# round 2, job ov2 (gpurun --gpus 2): the overlapped schedule under NCCL - bench configs[1] and configs[3] on 2 ranks, the sharded train_net entry
mkdir -p gpurun_out
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02ov2_bench_${N}gpu.json 2>gpurun_out/r02ov2_bench_${N}gpu.err; cut -c1-300 gpurun_out/r02ov2_bench_${N}gpu.json; tail -3 gpurun_out/r02ov2_bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --config 3 > gpurun_out/r02ov2_bench_cfg3_${N}gpu.json 2>gpurun_out/r02ov2_bench_cfg3_${N}gpu.err; cut -c1-300 gpurun_out/r02ov2_bench_cfg3_${N}gpu.json; tail -3 gpurun_out/r02ov2_bench_cfg3_${N}gpu.err
timeout 300 python -m pytest tests/test_gpu_entry.py -q --tb=short -x --timeout 200 2>&1 | tail -2
