#!/usr/bin/env python
"""bench.py - the driver's benchmark contract for the TTDG-MGM test-time-adaptation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of 8 synthetic 512x512 2-class images per GPU (BASELINE.json configs[1]) through BOTH passes of
adapteacher/engine/trainer.py:469-485:
  (1) test-time adaptation: detector forward in train mode (ResNet-50-FPN, RPN, box head) -> node sampler ->
      MGM3_unsup (attention adjacency, affinity, Sinkhorn 20 iters, GA-GM with on-device Hungarian, matching loss)
      -> backward through FPN + res3-res5 -> [NCCL all-reduce of the 26.97 M-element gradient bucket when N > 1]
      -> fused SGD step;
  (2) eval-mode inference with the adapted weights: detector forward, mask head, masks pasted to 512x512.
`value` = adapted images / s with the uint8 images resident in HBM; `e2e` = the same from pinned HOST images through
the plugin call (model(batched_inputs, branch='TTT') ... model(batched_inputs)) with the loss and a mask checksum read
back.  `roofline` = the Sinkhorn kernel of BASELINE.json configs[4] (N = 1024, 50 iterations), timed live.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ttdg-mgm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

IMAGES_PER_GPU = 8
IMG = 512
METRIC = "test_time_adapted_images_per_sec"
UNIT = "images/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------ workload
def make_inputs(rank):
    """Per-rank batch of seeded synthetic fundus-like images (uint8 3 x 512 x 512), as the dataset mapper
    would deliver them: list of dicts with 'image', 'height', 'width', 'image_id'."""
    from ttdg_b200 import synth
    out = []
    for i in range(IMAGES_PER_GPU):
        im = synth.fundus_like_image(rank * IMAGES_PER_GPU + i, IMG)
        out.append({"image": im["image"], "height": IMG, "width": IMG, "image_id": rank * IMAGES_PER_GPU + i})
    return out


def full_state():
    from ttdg_b200 import synth
    sd = dict(synth.detector_state_calibrated(0, 2))
    # matching head: the reference constructors' own init (affinity.py:33-42, mgm:124).  With it the adaptation is gentle
    # and the workload stays stationary over the run (100 detections / image, 30-45 nodes / graph); the "perturbed"
    # affinity used by some parity tests makes a RANDOM-init detector diverge within ~10 steps at lr 0.005.
    sd.update({"multi_matching_unsup." + k: v for k, v in synth.mgm_unsup_state(0).items()})
    sd["multi_matching_sup.U"] = synth.universe(0)
    return sd


def build_ours(device):
    from adapteacher.modeling.meta_arch.rcnn import DAobjTwoStagePseudoLabGeneralizedRCNN
    from ttdg_b200.optim import FlatSGD
    m = DAobjTwoStagePseudoLabGeneralizedRCNN(2).to(device)
    m.load_state_dict(full_state(), strict=False)
    opt = FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
    return m, opt


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ roofline leg
def sinkhorn_roofline(device, n=1024, batch=128, iters=50, launches=5, warm=3):
    """BASELINE.json configs[4]: batch x n x n fp32 (512 MiB > L2), 50 iterations, tau 0.05.  ALGORITHMIC bytes per
    launch = batch * n * n * 4 * 2 * iters (one read + one write of the matrix per half-iteration)."""
    from ttdg_b200 import ops
    s = torch.randn(batch, n, n, device=device)
    out = torch.empty_like(s)
    for _ in range(warm):
        ops.sinkhorn_stream(s, tau=0.05, max_iter=iters, out=out)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(launches)]
    for a, b in evs:
        a.record()
        ops.sinkhorn_stream(s, tau=0.05, max_iter=iters, out=out)
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / launches
    alg = batch * n * n * 4 * 2 * iters
    pk, how = peaks()
    achieved = alg / (ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_sinkhorn_stream_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    return {"bound": "hbm", "kernel": "sinkhorn_stream_kernel", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
            "peak_source": how + " (burst copy)", "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 3), "traffic": traffic,
            "algorithmic_bytes_per_launch": alg, "ms_per_launch": round(ms, 4),
            "workload": f"{batch} x {n} x {n} fp32, {iters} iterations, tau 0.05",
            "note": "matrix stays in distributed shared memory across iterations: DRAM traffic is ~2 passes, not 2*iters"}


# ------------------------------------------------------------------------------------------------ tensor roofline of the step
def conv_roofline(step_fn, steps=2):
    """The step's dominant kernel family (conv_tc_kernel + wgrad_tc_kernel, ~45 % of the device time): CUDA events around
    every launch of `steps` extra, untimed steps.  `achieved` = ALGORITHMIC flops (2 * pixels * Cin * Cout * taps: what an
    fp32 convolution needs) / summed kernel time; in the 3xTF32 parity mode the tensor pipe executes three
    TF32 MMAs per product, reported as `mma_tflops`.  `peak` = half of the measured dense bf16 rate (TF32 runs at half the
    16-bit rate; sustained figure: the kernels run inside a long step)."""
    from ttdg_b200 import _C
    lib = _C.lib()
    rec = []

    class Timed:
        def __init__(self, name, fn):
            self.name, self.fn = name, fn

        def __call__(self, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = self.fn(*a)
            e1.record()
            ints = [int(v) for v in a if isinstance(v, int) and not isinstance(v, bool)]
            rec.append((self.name, ints, e0, e1))
            return rc

    proxy = type("LibProxy", (), {})()
    for name in _C.SIGNATURES:
        fn = getattr(lib, name)
        setattr(proxy, name, Timed(name, fn) if name in ("ttdg_conv_tc", "ttdg_wgrad_tc", "ttdg_stem_tc") else fn)
    _C._lib = proxy
    try:
        for _ in range(steps):
            step_fn()
        torch.cuda.synchronize()
    finally:
        _C._lib = lib
    flops = ms = 0.0
    for name, a, e0, e1 in rec:
        if name == "ttdg_conv_tc":          # res_mode, relu, flip, N, H, W, Cin, Cout, R, S, pad, in_stride, ...
            N, H, W, Cin, Cout, R, S, pad, stride = a[3:12]
        elif name == "ttdg_wgrad_tc":       # precise, N, H, W, Cin, Cout, R, S, stride, pad
            N, H, W, Cin, Cout, R, S, stride, pad = a[1:10]
        else:                               # stem: Wp, relu, N, H, W  (7 x 7 stride 2, 3 -> 64)
            N, H, W = a[2:5]
            Cin, Cout, R, S, pad, stride = 3, 64, 7, 7, 3, 2
        Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        flops += 2.0 * N * Ho * Wo * Cin * Cout * R * S
        ms += e0.elapsed_time(e1)
    pk, how = peaks()
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / 2.0
    mult = 3 if os.environ.get("TTDG_CONV", "tf32x3") == "tf32x3" else 1
    achieved = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "conv_tc_kernel + wgrad_tc_kernel (tcgen05 kind::tf32)", "achieved": round(achieved, 1),
            "mma_tflops": round(achieved * mult, 1), "peak": round(peak, 1), "peak_source": how + " dense bf16 (sustained) / 2 = TF32",
            "unit": "TFLOP/s", "frac": round(achieved / peak, 3), "frac_mma": round(achieved * mult / peak, 3), "traffic": None,
            "launches_per_step": len(rec) // steps, "ms_per_step": round(ms / steps, 3),
            "algorithmic_tflop_per_step": round(flops / steps / 1e12, 3),
            "note": "3xTF32 parity mode: 3 TF32 MMAs per fp32-grade product; frac = algorithmic, frac_mma = tensor-pipe work"}


# ------------------------------------------------------------------------------------------------ CPU leg (oracle port)
def cpu_baseline(images_u8, steps):
    """The reference's algorithm on the host cores: oracle/ttt_port.Trainer (TTT step + eval pass) on `images_u8`."""
    from oracle import ttt_port                          # the one place bench.py executes oracle/: as the timed baseline
    from ttdg_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    tr = ttt_port.Trainer(synth.detector_state_calibrated(0, 2), synth.mgm_unsup_state(0), synth.universe(0))
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.ttt_step(images_u8)
        tr.eval_pass(images_u8)
    dt = (time.perf_counter() - t0) / steps
    return len(images_u8) / dt, dt


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = ("configs[1]: batch of 8 synthetic 512x512 2-class fundus-like images per GPU; per image one share of a "
                "test-time-adaptation step (Mask R-CNN R50-FPN fwd in train mode, node sampler, MGM3_unsup with Sinkhorn 20 "
                "iters + GA-GM, backward through FPN+res3-5, SGD) plus one eval forward with masks pasted at 512x512")
    config = {"workload": workload, "images_per_gpu": IMAGES_PER_GPU, "image_size": IMG, "num_classes": 2, "universe": 32,
              "sinkhorn_iters": 20,
              "conv_math": {"tf32x3": "tcgen05 kind::tf32 with hi/lo operand split + chunked TMEM accumulation (fp32-grade, 2e-6 per layer)",
                            "tf32": "tcgen05 single-pass TF32", "simt": "fp32 CUDA-core FMA"}[os.environ.get("TTDG_CONV", "tf32x3")],
              "weights": "random init, FrozenBN statistics calibrated on synthetic images (no checkpoint offline)",
              "parallelism": f"image-sharded x{world}, NCCL all-reduce of the gradient bucket",
              "l2": "flushed between timed steps (256 MiB memset outside the per-step event pairs)"}

    if args.impl == "reference":
        if rank != 0:
            return
        # The reference's own Python cannot travel to the GPU box (/root/reference is absent there, Detectron2 0.5 is not
        # installable, and pure Python cannot be compiled into oracle/_ref): the oracle port stands in (kind "port").
        n_img = 2                                            # bounded sample: 2 of the 8 images per step
        images = [d["image"] for d in make_inputs(0)[:n_img]]
        steps = max(1, min(args.steps, 2))
        cpu_baseline(images, 1)                              # warm-up
        val, dt = cpu_baseline(images, steps)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
                          "steps": steps, "warmup": 1, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                           "sample": f"{steps} steps of {n_img} images (TTT step + eval pass), torch CPU, all cores"},
                          "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner to stdout at any debug level >= VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # an explicit INFO / TRACE request goes to stderr
        torch.distributed.init_process_group("nccl", device_id=device)
    from ttdg_b200 import _C
    lib = _C.lib()
    m, opt = build_ours(device)
    inputs_host = make_inputs(rank)
    for d in inputs_host:
        d["image"] = d["image"].pin_memory()
    inputs_dev = [dict(d, image=d["image"].to(device)) for d in inputs_host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    stats = {"skipped": 0}
    host_res = torch.empty(IMAGES_PER_GPU + 1, dtype=torch.float64).pin_memory()

    def step(inputs, readback):
        m.train()                                            # pass 1: adaptation (trainer.py:469-482)
        loss, _, _, _ = m(inputs, branch="TTT")
        if loss is None:
            stats["skipped"] += 1
        else:
            opt.zero_grad()
            loss.backward()
            opt.step(world)
        m.eval()                                             # pass 2: inference with the adapted weights (trainer.py:484-485)
        out = m(inputs)
        if readback:                                         # one D2H read of the step's result: per-image mask pixel counts + the loss
            res = torch.stack([o["instances"].pred_masks.sum().to(torch.float64) for o in out] +
                              [loss.detach().to(torch.float64) if loss is not None else torch.full((), float("nan"), dtype=torch.float64, device=device)])
            host_res.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return (float(host_res[-1]) if loss is not None else None), int(host_res[:-1].sum())
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    for _ in range(max(args.warmup, 3)):
        step(inputs_dev, False)
    barrier()
    clocks = ClockSampler(local_rank)
    l0 = lib.ttdg_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        step(inputs_dev, False)
        b.record()
    barrier()
    launches = lib.ttdg_launch_count() - l0
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    value = IMAGES_PER_GPU * world * args.steps / (ms_total * 1e-3)

    # ---- end to end through the plugin call with pinned HOST images; loss + mask checksum read back every step
    h2d = sum(d["image"].numel() for d in inputs_host)
    for _ in range(2):
        step(inputs_host, True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        last = step(inputs_host, True)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_val = IMAGES_PER_GPU * world * args.steps / (float(t.item()) * 1e-3)
    clk = clocks.stop()

    roof_conv = conv_roofline(lambda: step(inputs_dev, False))           # every rank runs it (the all-reduce inside is collective)
    if rank == 0:
        roof = sinkhorn_roofline(device)
        n_cpu = 2
        cpu_val, cpu_dt = (cpu_baseline([d["image"] for d in make_inputs(0)[:n_cpu]], 1) if world == 1 else (None, None))
        aux = m.multi_matching_unsup.last_aux
        info = aux["info"].cpu().tolist()
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * (IMAGES_PER_GPU + 1),
                        "call": "model(batched_inputs, branch='TTT') + backward + FlatSGD.step + model(batched_inputs) from pinned host images",
                        "last_loss": last[0], "mask_pixels": last[1]},
                "gpu_launches": int(launches), "skipped_steps": stats["skipped"],
                "gagm": {"iterations": info[0], "lap_calls": info[3], "graphs": len(aux["sizes"]), "nodes": int(sum(aux["sizes"]))},
                "clocks": clk, "roofline": roof, "roofline_step_dominant": roof_conv,
                "cpu_baseline": ({"value": round(cpu_val, 4), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                  "sample": f"1 step of {n_cpu} of the 8 images (TTT step + eval pass) with the oracle port on torch "
                                            f"CPU, {cpu_dt:.1f} s"} if world == 1 else None)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
