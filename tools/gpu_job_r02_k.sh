# round 2, job k (multi-GPU): weak scaling with the bucketed, overlapped gradient all-reduce
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02k_bench_${N}gpu.json 2>gpurun_out/r02k_bench_${N}gpu.err; cut -c1-300 gpurun_out/r02k_bench_${N}gpu.json; tail -3 gpurun_out/r02k_bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --config 3 > gpurun_out/r02k_bench_cfg3_${N}gpu.json 2>gpurun_out/r02k_bench_cfg3_${N}gpu.err; cut -c1-300 gpurun_out/r02k_bench_cfg3_${N}gpu.json; tail -3 gpurun_out/r02k_bench_cfg3_${N}gpu.err
