# 2-GPU check of both bench arms and of the sharded train_net entry (gpurun --gpus 2)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2>gpurun_out/bench_2gpu.err; echo "rc=$?"; wc -l gpurun_out/bench_2gpu.json
python -c "
import json;d=json.load(open('gpurun_out/bench_2gpu.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['gagm'],d['roofline_step_dominant']['achieved'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-160
python - <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'ttdg-mgm_b200')
import bench
torch.save({"model": bench.full_state()}, "gpurun_out/model_final.pth")
PY
timeout 600 python ttdg-mgm_b200/train_net.py --eval-only --num-gpus 2 --config ttdg-mgm_b200/configs/test_segment_synthetic.yaml MODEL.WEIGHTS gpurun_out/model_final.pth OUTPUT_DIR gpurun_out/out2 DATASETS.TEST '("synthetic_fundus_18_256",)' TEST.BATCH 4 2>&1 | tail -2 | cut -c1-300
rm -f gpurun_out/model_final.pth; rm -rf gpurun_out/out2/*.pth
