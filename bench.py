#!/usr/bin/env python
"""bench.py - the driver's benchmark contract for the TTDG-MGM test-time-adaptation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = this rank's shard of 8 synthetic images through BOTH passes of adapteacher/engine/trainer.py:469-485, batched
as the reference's test loader batches them (TEST.BATCH, drop_last False):
  (1) test-time adaptation, per batch: detector forward in train mode (ResNet-50-FPN, RPN, box head) -> node sampler ->
      MGM3_unsup (attention adjacency, affinity, Sinkhorn 20 iters, GA-GM with on-device Hungarian, matching loss)
      -> backward through FPN + res3-res5 -> [NCCL all-reduce of the 26.97 M-element gradient bucket when N > 1]
      -> fused SGD step;
  (2) eval-mode inference with the adapted weights, per batch: detector forward, mask head, masks pasted at image size.
--config 1 (default) = BASELINE.json configs[1]: 8 x 512 x 512, 2 classes, TEST.BATCH 8, fp32-grade convolutions;
--config 2 = configs[2]: the same with the bf16 backbone (fp32 matching stage), meant for --gpus 4;
--config 3 = configs[3]: 8 x 384 x 384 polyp-like, 1 class, TEST.BATCH 5 (matching problems of 5 + 3 graphs), meant for --gpus 8.
Schedule (default, TTDG_OVERLAP=0 switches it off): a step is one mini-dataset; pass (2) of step i runs from a weight snapshot on
a second stream inside the GA-GM solver windows of step i + 1's pass (1) - the product's schedule for consecutive datasets
(adapteacher/engine/trainer.py OverlappedEval; identical results, tests/test_gpu_overlap.py).  The timed region starts with nothing
pending and ends with the last pass (2) drained: K steps = K adaptation passes + K evaluation passes, all work inside.
`value` = adapted images / s with the uint8 images resident in HBM; `e2e` = the same from pinned HOST images through the
plugin call (model(batched_inputs, branch='TTT') ... model(batched_inputs)) with the loss and a mask checksum read back.
`roofline` = the step's dominant kernel family (tcgen05 convolutions), timed live; `sinkhorn_microbench` = configs[4].
"""
import argparse
import contextlib
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
METRIC = "test_time_adapted_images_per_sec"
UNIT = "images/s"
CONFIGS = {
    1: dict(name="configs[1]", size=512, test_batch=8, num_classes=2, polyp=False, conv="tf32x3", dtype="f32", first=0),
    2: dict(name="configs[2]", size=512, test_batch=8, num_classes=2, polyp=False, conv="bf16", dtype="bf16", first=0),
    3: dict(name="configs[3]", size=384, test_batch=5, num_classes=1, polyp=True, conv="tf32x3", dtype="f32", first=500),
}
CONV_MATH = {"tf32x3": "tcgen05 kind::tf32 with hi/lo operand split + chunked TMEM accumulation (fp32-grade)",
             "tf32": "tcgen05 single-pass TF32", "simt": "fp32 CUDA-core FMA",
             "bf16": "tcgen05 kind::f16 on bf16 activations / weights, fp32 TMEM accumulation, fp32 master weights"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------ workload
def make_batches(rank, cfg):
    """This rank's shard of seeded synthetic images (uint8 3 x S x S) as the dataset mapper would deliver them -
    dicts with 'image', 'height', 'width', 'image_id' - cut into TEST.BATCH-sized batches (data/build.py:141-146)."""
    from ttdg_b200 import synth
    S = cfg["size"]
    items = []
    for i in range(IMAGES_PER_GPU):
        idx = cfg["first"] + rank * IMAGES_PER_GPU + i
        im = synth.fundus_like_image(idx, S, cfg["polyp"])
        items.append({"image": im["image"], "height": S, "width": S, "image_id": idx})
    tb = cfg["test_batch"]
    return [items[i:i + tb] for i in range(0, len(items), tb)]


def full_state(cfg):
    from ttdg_b200 import synth
    sd = dict(synth.detector_state_calibrated(0, cfg["num_classes"]))
    # matching head: the reference constructors' own init (affinity.py:33-42, mgm:124).  With it the adaptation is gentle
    # and the workload stays stationary over the run (100 detections / image, 30-45 nodes / graph); the "perturbed"
    # affinity used by some parity tests makes a RANDOM-init detector diverge within ~10 steps at lr 0.005.
    sd.update({"multi_matching_unsup." + k: v for k, v in synth.mgm_unsup_state(0).items()})
    sd["multi_matching_sup.U"] = synth.universe(0)
    return sd


OVERLAP = [os.environ.get("TTDG_OVERLAP", "1") != "0"]     # the trainer's schedule (TEST.OVERLAP_EVAL); 0 = strictly sequential


def build_ours(device, cfg=None):
    from adapteacher.modeling.meta_arch.rcnn import DAobjTwoStagePseudoLabGeneralizedRCNN
    from ttdg_b200.optim import FlatSGD
    cfg = cfg or CONFIGS[1]
    m = DAobjTwoStagePseudoLabGeneralizedRCNN(cfg["num_classes"]).to(device)
    m.load_state_dict(full_state(cfg), strict=False)
    opt = FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4,
                  buckets=[len(g) for g in m.adapted_parameter_groups()], on_step=m.refresh_weight_copies)
    return m, opt


def make_inputs(rank, cfg=None):          # (tools/run_kernels.py): one flat batch of configs[1]
    return [d for b in make_batches(rank, cfg or CONFIGS[1]) for d in b]


def describe(cfg, world, impl):
    S = cfg["size"]
    tb = cfg["test_batch"]
    sizes = [min(tb, IMAGES_PER_GPU - i) for i in range(0, IMAGES_PER_GPU, tb)]
    d = {"workload": "%s: %d synthetic %dx%d %s images per GPU, TEST.BATCH %d (matching problems of %s graphs); per image one "
                     "share of a test-time-adaptation step (Mask R-CNN R50-FPN fwd in train mode, node sampler, MGM3_unsup with "
                     "Sinkhorn 20 iters + GA-GM, backward through FPN+res3-5, SGD) plus one eval forward with masks pasted at "
                     "%dx%d" % (cfg["name"], IMAGES_PER_GPU, S, S, "polyp-like 1-class" if cfg["polyp"] else "fundus-like 2-class",
                                tb, "+".join(map(str, sizes)), S, S),
         "images_per_gpu": IMAGES_PER_GPU, "image_size": S, "num_classes": cfg["num_classes"], "test_batch": tb,
         "universe": 32, "sinkhorn_iters": 20,
         "weights": "random init, FrozenBN statistics calibrated on synthetic images (no checkpoint offline)"}
    if impl == "ours":
        d["conv_math"] = CONV_MATH[cfg["conv"]]
        d["parallelism"] = (f"image-sharded x{world}; per adaptation step the flat gradient buffer is all-reduced over NCCL in 3 buckets "
                            "(affinity + FPN + res5 | res4 | res3) started from the backward pass as each completes")
        d["l2"] = "flushed before every timed step (256 MiB memset inside the timed region)"
        d["schedule"] = ("a step = one mini-dataset of the rank's images: pass 1 (adaptation over its batches), then pass 2 (evaluation "
                         "with the adapted weights), as trainer.py:469-485 orders them per dataset" +
                         ("; pass 2 of step i runs on a second stream inside the GA-GM solver windows of step i + 1's pass 1 from a weight "
                          "snapshot (adapteacher.engine.trainer.OverlappedEval, the product's schedule for consecutive datasets): the "
                          "timed region starts with nothing pending and ends with the last pass 2 drained, so K steps hold exactly K "
                          "adaptation passes and K evaluation passes" if OVERLAP[0] else "; sequential (TTDG_OVERLAP=0)"))
    else:
        d["implementation"] = "oracle/ttt_port.Trainer: the reference's algorithm restated on torch CPU (fp32), all host threads"
    return d


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


# ------------------------------------------------------------------------------------------------ configs[4] microbench
def sinkhorn_microbench(device, n=1024, batch=128, iters=50, launches=5, warm=3):
    """BASELINE.json configs[4]: batch x n x n fp32 (512 MiB > L2), 50 iterations, tau 0.05.  ALGORITHMIC bytes per
    launch = batch * n * n * 4 * 2 * iters (one read + one write of the matrix per half-iteration).  The kernel
    keeps the matrix in distributed shared memory, so against HBM it is not a roofline (frac > 1): the honest bound is on
    chip - see `onchip`."""
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
    for name in ("r02_sinkhorn_stream_traffic.json", "r01_sinkhorn_stream_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            break
        except Exception:
            pass
    # on-chip bound: every element is touched once per half-iteration from shared memory (4 B) and goes through one ex2;
    # 148 SMs x 128 B/clk of shared-memory bandwidth and 148 x 4 x 16 MUFU lanes / clk at the sampled SM clock
    elems = batch * n * n * iters
    return {"bound": "hbm", "kernel": "sinkhorn_stream_kernel", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"],
            "peak_source": how + " (burst copy)", "unit": "GB/s", "frac": round(achieved / pk["hbm_gbs"], 3), "traffic": traffic,
            "algorithmic_bytes_per_launch": alg, "ms_per_launch": round(ms, 4),
            "workload": f"{batch} x {n} x {n} fp32, {iters} iterations, tau 0.05",
            "onchip": {"elements_per_launch": elems, "gelem_per_s": round(elems / (ms * 1e-3) / 1e9, 1),
                       "mufu_peak_gelem_per_s": round(148 * 64 * 1.965, 1), "frac_mufu": round(elems / (ms * 1e-3) / 1e9 / (148 * 64 * 1.965), 3),
                       "note": "one ex2 + one shared-memory read per element per half-iteration; MUFU peak = 148 SMs x 64 lanes / clk at 1965 MHz"},
            "note": "matrix stays in distributed shared memory across iterations: DRAM traffic is ~2 passes, not 2*iters"}


# ------------------------------------------------------------------------------------------------ live per-launch timing
@contextlib.contextmanager
def timed_lib(names, lead_cycles=0):
    """Wraps the named C-ABI entry points with CUDA events (torch's current stream = the launch stream); yields the records.
    lead_cycles > 0: a device-side spin of that many cycles is enqueued before the first event of every timed call, so that the
    launch is already queued when the first event is stamped - the pair then brackets the kernel alone and not the host's launch
    latency (~4 us per call when the stream has run dry, ~1 ms over the 263 conv launches of a step)."""
    from ttdg_b200 import _C
    lib = _C.lib()
    rec = []

    class Timed:
        def __init__(self, name, fn):
            self.name, self.fn = name, fn

        def __call__(self, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if lead_cycles:
                torch.cuda._sleep(lead_cycles)
            e0.record()
            rc = self.fn(*a)
            e1.record()
            rec.append((self.name, [int(v) for v in a if isinstance(v, int) and not isinstance(v, bool)], e0, e1))
            return rc

    proxy = type("LibProxy", (), {})()
    for name in _C.SIGNATURES:
        fn = getattr(lib, name)
        setattr(proxy, name, Timed(name, fn) if name in names else fn)
    _C._lib = proxy
    try:
        yield rec
    finally:
        _C._lib = lib


def conv_roofline(step_fn, conv, steps=2):
    """The step's dominant kernel family (conv_tc_kernel + wgrad_tc_kernel): CUDA events around every launch of `steps`
    extra, untimed steps.  `achieved` = ALGORITHMIC flops (2 * pixels * Cin * Cout * taps: what the convolution needs,
    SURVEY 8d) / summed kernel time; in the 3xTF32 parity mode the tensor pipe executes three TF32 MMAs per product, reported
    as `mma_tflops`.  `peak` = the measured dense bf16 rate (sustained: the kernels run inside a long step), halved for the
    TF32 modes (TF32 runs at half the 16-bit rate)."""
    with timed_lib(("ttdg_conv_tc", "ttdg_wgrad_tc", "ttdg_stem_tc", "ttdg_stem_tc2", "ttdg_conv_tc_bf16"), lead_cycles=100000) as rec:
        for _ in range(steps):
            step_fn()
        torch.cuda.synchronize()
    flops = ms = 0.0
    for name, a, e0, e1 in rec:
        if name == "ttdg_conv_tc":          # res_mode, relu, flip, N, H, W, Cin, Cout, R, S, pad, in_stride, ...
            N, H, W, Cin, Cout, R, S, pad, stride = a[3:12]
        elif name == "ttdg_conv_tc_bf16":   # res_bf16, res_mode, relu, flip, N, H, W, Cin, Cout, R, S, pad, in_stride, ...
            N, H, W, Cin, Cout, R, S, pad, stride = a[4:13]
        elif name == "ttdg_wgrad_tc":       # precise, N, H, W, Cin, Cout, R, S, stride, pad
            N, H, W, Cin, Cout, R, S, stride, pad = a[1:10]
        else:                               # stem: Wp, relu, N, H, W  (7 x 7 stride 2, 3 -> 64)
            N, H, W = a[2:5]
            Cin, Cout, R, S, pad, stride = 3, 64, 7, 7, 3, 2
        Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        flops += 2.0 * N * Ho * Wo * Cin * Cout * R * S
        ms += e0.elapsed_time(e1)
    pk, how = peaks()
    full = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    peak = full if conv == "bf16" else full / 2.0
    mult = 3 if conv == "tf32x3" else 1
    achieved = flops / (ms * 1e-3) / 1e12
    kind = "kind::f16 (bf16)" if conv == "bf16" else "kind::tf32"
    traffic = None                                          # DRAM bytes per launch of the family, from one ncu --set full capture
    if conv == "tf32x3":
        try:
            with open(os.path.join(ROOT, "profiles", "r02_conv_tc_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            traffic = None
    return {"bound": "tensor", "kernel": "conv_tc_kernel + wgrad_tc_kernel (tcgen05 %s)" % kind, "achieved": round(achieved, 1),
            "mma_tflops": round(achieved * mult, 1), "peak": round(peak, 1),
            "peak_source": how + (" dense bf16 (sustained)" if conv == "bf16" else " dense bf16 (sustained) / 2 = TF32"),
            "unit": "TFLOP/s", "frac": round(achieved / peak, 3), "frac_mma": round(achieved * mult / peak, 3), "traffic": traffic,
            "launches_per_step": len(rec) // steps, "ms_per_step": round(ms / steps, 3),
            "algorithmic_tflop_per_step": round(flops / steps / 1e12, 3),
            "note": ("3xTF32 parity mode: 3 TF32 MMAs per fp32-grade product; frac = algorithmic, frac_mma = tensor-pipe work"
                     if conv == "tf32x3" else "frac = algorithmic flops / measured dense peak") +
                    "; CUDA events around every launch, each preceded by a ~50 us device-side spin so that the pair brackets the kernel only"}


# ------------------------------------------------------------------------------------------------ CPU leg (oracle port)
def cpu_baseline(cfg, batches_u8, steps):
    """The reference's algorithm on the host cores: oracle/ttt_port.Trainer (adaptation pass over the batches, then the eval
    pass over the same batches)."""
    from oracle import ttt_port                          # the one place bench.py executes oracle/: as the timed baseline
    from ttdg_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    tr = ttt_port.Trainer(synth.detector_state_calibrated(0, cfg["num_classes"]), synth.mgm_unsup_state(0), synth.universe(0))
    n = sum(len(b) for b in batches_u8)
    t0 = time.perf_counter()
    for _ in range(steps):
        for b in batches_u8:
            tr.ttt_step(b)
        for b in batches_u8:
            tr.eval_pass(b)
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt


def parity_sample(cfg):
    """Measured parity of the CUDA path against the oracle on a 2-image sample of this config (tests/_parity.py; the full
    8-image versions are tests/test_gpu_parity_configs.py).  Checker use of oracle/, inside the cpu_baseline leg."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _parity
    pc = dict(size=cfg["size"], batch=2, num_classes=cfg["num_classes"], polyp=cfg["polyp"], first=cfg["first"])
    m, sd_det, _, _ = _parity.build_model(cfg["num_classes"])
    ev = _parity.eval_parity(m, sd_det, pc, with_f64=False, log=lambda *a: None)
    fr, mb = ev["free_running_gpu_vs_fp32"], ev["mask_branch_forced_detections"]
    return {"sample": "eval pass on 2 images of this config, CUDA path vs oracle (fp32 restatement)",
            "matched_frac": round(fr["matched_frac"], 4), "miou_delta_matched": fr["miou_delta_matched"],
            "miou_delta_all": fr["miou_delta_all"], "miou_delta_forced_detections": mb["miou_delta"],
            "pyramid_rel_max": ev["pyramid_rel_max"]}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        # The reference's own Python cannot travel to the GPU box (/root/reference is absent there, Detectron2 0.5 is not
        # installable, and pure Python cannot be compiled into oracle/_ref): the oracle port stands in (kind "port").  It runs
        # the SAME per-GPU workload as our arm: all 8 images of rank 0's shard, both passes, every step.
        batches = [[d["image"] for d in b] for b in make_batches(0, cfg)]
        steps = max(2, min(args.steps, 3))
        warm = 1
        cpu_baseline(cfg, batches, warm)
        val, dt = cpu_baseline(cfg, batches, steps)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
                          "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": describe(cfg, 1, "reference"),
                          "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                           "sample": f"{steps} steps of all {IMAGES_PER_GPU} images of one GPU's shard (adaptation pass + "
                                                     f"eval pass), torch CPU, all cores"},
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
    from ttdg_b200 import _C, detector
    lib = _C.lib()
    conv = os.environ.get("TTDG_CONV", cfg["conv"])
    detector.set_conv_mode(conv)
    cfg = dict(cfg, conv=conv)
    m, opt = build_ours(device, cfg)
    if world > 1:                                            # gradient buckets are all-reduced while the backward still runs
        opt.enable_overlap(world)
        detector.GRAD_READY_HOOK[0] = opt.grad_ready
    batches_host = make_batches(rank, cfg)
    for b in batches_host:
        for d in b:
            d["image"] = d["image"].pin_memory()
    batches_dev = [[dict(d, image=d["image"].to(device)) for d in b] for b in batches_host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    stats = {"skipped": 0}
    host_res = torch.empty(IMAGES_PER_GPU + 1, dtype=torch.float64).pin_memory()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def pass1(batches, marks=None):
        loss = None
        m.train()                                            # pass 1: adaptation (trainer.py:469-482)
        for inputs in batches:
            t = [ev()] if marks is not None else None
            loss, _, _, _ = m(inputs, branch="TTT")
            if loss is None:
                stats["skipped"] += 1
                continue
            opt.zero_grad()
            if t is not None:
                t.append(ev())
            loss.backward()
            if t is not None:
                t.append(ev())
            opt.allreduce(world)
            if t is not None:
                t.append(ev())
            opt.step(world, reduce=False)
            if t is not None:
                t.append(ev())
                marks.append(t)
        return loss

    def read_back(out, loss):                                # one D2H read of a step's result: per-image mask pixel counts + the loss
        res = torch.stack([o["instances"].pred_masks.sum().to(torch.float64) for o in out] +
                          [loss.detach().to(torch.float64) if loss is not None else torch.full((), float("nan"), dtype=torch.float64, device=device)])
        host_res.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return (float(host_res[-1]) if loss is not None else None), int(host_res[:-1].sum())

    def step(batches, readback, marks=None):
        """Strictly sequential: pass 1, then pass 2 with the adapted weights (instrumented steps, per-kernel timing, TTDG_OVERLAP=0)."""
        loss = pass1(batches, marks)
        m.eval()                                             # pass 2: inference with the adapted weights (trainer.py:484-485)
        out = []
        for inputs in batches:
            out += m(inputs)
        return read_back(out, loss) if readback else None

    # ---- the product's schedule (adapteacher/engine/trainer.py): pass 2 of a step inside the solver windows of the next step's pass 1
    from adapteacher.engine.trainer import OverlappedEval
    from ttdg_b200 import ops

    class Sink:                                              # the evaluator of the bench: keeps (e2e: reads back) the step's masks
        def __init__(self):
            self.readback, self.loss, self.last, self.out = False, None, None, []

        def reset(self):
            self.out = []

        def process(self, inputs, outputs):
            self.out += outputs

        def evaluate(self):
            if self.readback:                                # on the evaluation stream: waits for that stream only
                self.last = read_back(self.out, self.loss)
            self.out = []
            return None

    pipe, sink = OverlappedEval(m), Sink()

    def drain():
        if pipe.active:
            with torch.cuda.stream(pipe.stream):
                pipe.drain()

    def step_overlapped(batches, readback):
        ops.SOLVER_WINDOW_HOOK[0] = pipe.window if pipe.active else None
        try:
            loss = pass1(batches)
        finally:
            ops.SOLVER_WINDOW_HOOK[0] = None
        drain()                                              # what the windows left of the previous step's pass 2 (+ its read-back)
        sink.readback, sink.loss = readback, loss
        pipe.begin("bench", batches, sink)                   # snapshot of the adapted weights; evaluated during the next step
        return sink.last

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    if world > 1:
        # one process per GPU shares the host: keep every rank's launch thread on its own cores (the step issues ~560 launches
        # from Python; a rank that loses its core for a few ms makes all the others wait in the gradient all-reduce)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // world
            if per >= 1:
                os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
        except (AttributeError, OSError):
            pass
    run_step = step_overlapped if OVERLAP[0] else step
    warm = max(args.warmup, 3)
    for _ in range(warm):
        run_step(batches_dev, False)
    drain()                                                  # nothing pending when the timed region starts
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None      # one nvidia-smi poller per job, not per rank
    l0 = lib.ttdg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                        # L2 flush between steps (inside the timed region: 256 MiB memset)
        run_step(batches_dev, False)
    drain()                                                  # the last step's pass 2 belongs to the timed region
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps  # host time to ENQUEUE a step (no sync inside)
    e1.record()
    barrier()
    launches = lib.ttdg_launch_count() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = IMAGES_PER_GPU * world * args.steps / (ms_total * 1e-3)

    # ---- end to end through the plugin call with pinned HOST images; loss + mask checksum read back every step
    h2d = sum(d["image"].numel() for b in batches_host for d in b)
    for _ in range(2):
        run_step(batches_host, True)
    drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        flush.zero_()
        last = run_step(batches_host, True)
    if OVERLAP[0]:
        drain()
        last = sink.last
    e1.record()
    barrier()
    windows = pipe.windows
    e2e_val = IMAGES_PER_GPU * world * args.steps / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    clk = clocks.stop() if clocks is not None else None

    # ---- where the time goes, per rank (3 extra instrumented steps: CUDA events at the stage boundaries + around the solver)
    marks = []
    barrier()                                                # ranks enter the instrumented steps together (allreduce_wait = skew of the step only)
    with timed_lib(("ttdg_gagm_solve",)) as grec:
        for _ in range(3):
            flush.zero_()
            step(batches_dev, False, marks)
        torch.cuda.synchronize()
    n3 = 3.0
    mine = [sum(t[0].elapsed_time(t[1]) for t in marks) / n3, sum(t[1].elapsed_time(t[2]) for t in marks) / n3,
            sum(t[2].elapsed_time(t[3]) for t in marks) / n3, sum(t[3].elapsed_time(t[4]) for t in marks) / n3,
            sum(a.elapsed_time(b) for _, _, a, b in grec) / n3, host_ms]
    tl = torch.tensor(mine, dtype=torch.float64, device=device)
    if world > 1:
        allr = [torch.zeros_like(tl) for _ in range(world)]
        torch.distributed.all_gather(allr, tl)
        allr = torch.stack(allr).cpu()
    else:
        allr = tl.cpu().unsqueeze(0)
    keys = ("ttt_forward_ms", "backward_ms", "allreduce_wait_ms", "sgd_ms", "gagm_ms", "host_enqueue_ms")
    per_rank = {k: {"max": round(float(allr[:, i].max()), 3), "min": round(float(allr[:, i].min()), 3)} for i, k in enumerate(keys)}
    per_rank["note"] = ("per step, mean of 3 instrumented steps, max / min over ranks; gagm_ms is part of ttt_forward_ms; "
                        "host_enqueue_ms = host time to issue one step of the timed loop (no sync inside)")

    roof_conv = conv_roofline(lambda: step(batches_dev, False), conv)     # every rank runs it (the all-reduce inside is collective)
    if rank == 0:
        micro = sinkhorn_microbench(device)
        cpu = par = None
        if world == 1:
            cpu_val, cpu_dt = cpu_baseline(cfg, [[d["image"] for d in b] for b in make_batches(0, cfg)], 1)
            cpu = {"value": round(cpu_val, 4), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"1 step of all {IMAGES_PER_GPU} images (adaptation pass + eval pass) with the oracle port on torch CPU, {cpu_dt:.1f} s"}
            try:
                par = parity_sample(cfg)
            except Exception as e:                          # the parity sample must never cost the bench line
                par = {"error": repr(e)[:200]}
        aux = m.multi_matching_unsup.last_aux
        info = aux["info"].cpu().tolist()
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"] if conv == cfg["conv"] else "f32", "data": "synthetic",
                "config": describe(cfg, world, "ours"),
                "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * (IMAGES_PER_GPU + 1),
                        "call": ("model(batched_inputs, branch='TTT') + backward + FlatSGD.step per batch, then model(batched_inputs) per batch, "
                                 "from pinned host images" + ("; the second pass runs through OverlappedEval (weight snapshot, second stream) inside the "
                                                              "next step's solver windows, its result is read back there" if OVERLAP[0] else "")),
                        "last_loss": last[0], "mask_pixels": last[1]},
                "gpu_launches": int(launches), "skipped_steps": stats["skipped"],
                "overlap": ({"pass2_batches_in_solver_windows": int(windows), "note": "all timed + warm-up loops of this run"} if OVERLAP[0] else None),
                "gagm": {"iterations": info[0], "lap_calls": info[3], "lap_fallbacks_graph0": info[7], "graphs": len(aux["sizes"]),
                         "nodes": int(sum(aux["sizes"])), "ms": per_rank["gagm_ms"], "lap_row_relaxations_graph0": info[5],
                         "cta0_kcycles": {"kernel": info[8], "hungarian_stage": info[9], "in_lap": info[10], "barrier_wait": info[11]}},
                "per_rank": per_rank, "clocks": clk, "roofline": roof_conv, "sinkhorn_microbench": micro,
                "cpu_baseline": cpu, "parity": par}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
