"""Launches individual hot kernels a few times - the target command for `ncu` captures (see profiles/README.md).
    python tools/run_kernels.py sinkhorn_stream | gagm | mgm_step"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ttdg-mgm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from ttdg_b200 import ops, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "sinkhorn_stream"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
if what == "sinkhorn_stream":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    batch = (512 << 20) // (n * n * 4)
    s = torch.randn(batch, n, n, device=dev)
    out = torch.empty_like(s)
    for _ in range(reps):
        ops.sinkhorn_stream(s, tau=0.05, max_iter=50, out=out)
elif what in ("gagm", "mgm_step"):
    from adapteacher.modeling.GModule.multi_graph_matching import MGM3_unsup
    m = MGM3_unsup(2, 32).to(dev)
    m.load_state_dict(synth.perturb_affinity_state(synth.mgm_unsup_state(0), 0))
    m.train()
    sizes = (33, 34, 33, 33, 34, 33, 34, 33)
    nodes, labels, _ = synth.mgm_inputs(sizes, 77)
    U = synth.universe(0).to(dev)
    for _ in range(reps):
        loss = m([n.to(dev).requires_grad_(True) for n in nodes], [l.to(dev) for l in labels], U)
        loss.backward()
    print("gagm info", m.last_aux["info"].tolist())
elif what == "full_step":
    sys.path.insert(0, ROOT)
    import bench
    m, opt = bench.build_ours(dev)
    inputs = [dict(d, image=d["image"].to(dev)) for d in bench.make_inputs(0)]
    for r in range(reps):
        if r == reps - 1:                                  # ncu --profile-from-start off: exactly the last step is captured
            torch.cuda.synchronize(); torch.cuda.profiler.start()
        m.train()
        loss, _, _, _ = m(inputs, branch="TTT")
        opt.zero_grad(); loss.backward(); opt.step(1)
        m.eval()
        out = m(inputs)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    print("loss", float(loss), "dets", [len(o["instances"]) for o in out])
elif what == "timing":
    # CUDA-event timing of the stages of one full step (not under a profiler)
    sys.path.insert(0, ROOT)
    import bench
    from ttdg_b200.structures import Boxes, Instances
    m, opt = bench.build_ours(dev)
    det = m._det[0]
    inputs = [dict(d, image=d["image"].to(dev)) for d in bench.make_inputs(0)]
    images = [d["image"] for d in inputs]
    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    import collections
    acc = collections.defaultdict(float)
    for it in range(reps + 2):
        m.train()
        t0 = ev()
        feats = det.features(images); t1 = ev()
        props = det.proposal_generator.predict(feats, (512, 512), training=True); t2 = ev()
        dets = det.roi_heads.forward_box(feats, props, (512, 512)); t3 = ev()
        insts = [Instances((512, 512), pred_boxes=Boxes(b), scores=s, pred_classes=c) for b, s, c in dets]
        nodes, labels = m.graph_generator([f.permute(0, 3, 1, 2) for f in feats], insts); t4 = ev()
        loss = m.multi_matching_unsup(nodes, labels, m.multi_matching_sup.U); t5 = ev()
        opt.zero_grad(); loss.backward(); t6 = ev()
        opt.step(1); t7 = ev()
        m.eval()
        with torch.no_grad():                                    # = m(inputs), stage by stage
            e_feats = det.features(images); t8 = ev()
            e_props = det.proposal_generator.predict(e_feats, (512, 512), training=False); t9 = ev()
            e_dets = det.roi_heads.forward_box(e_feats, e_props, (512, 512)); t10 = ev()
            out = det.roi_heads.forward_mask(e_feats, e_dets, (512, 512), (512, 512)); t11 = ev()
        torch.cuda.synchronize()
        if it >= 2:
            for k, a, b in (("backbone_fwd", t0, t1), ("rpn", t1, t2), ("box_head", t2, t3), ("sampler", t3, t4), ("mgm_fwd", t4, t5),
                            ("backward", t5, t6), ("sgd", t6, t7), ("eval_backbone", t7, t8), ("eval_rpn", t8, t9), ("eval_box", t9, t10),
                            ("eval_mask", t10, t11), ("TOTAL", t0, t11)):
                acc[k] += a.elapsed_time(b) / reps
    info = m.multi_matching_unsup.last_aux["info"].tolist()
    print({k: round(v, 2) for k, v in acc.items()}, "gagm iters", info[0], "graph-0 LAP steps", info[5], "hops", info[6])
elif what == "layers":
    # per-call device time of every library launch inside a real full step (hot caches, CUDA events around each call),
    # grouped by entry point + geometry; conv rows carry their fp32-equivalent TFLOP/s (2 * pixels * Cin * Cout * taps)
    sys.path.insert(0, ROOT)
    import collections
    import bench
    from ttdg_b200 import _C
    m, opt = bench.build_ours(dev)
    inputs = [dict(d, image=d["image"].to(dev)) for d in bench.make_inputs(0)]
    L = _C.lib()
    rec, on = [], [False]
    class Timed:
        def __init__(self, name, fn):
            self.name, self.fn = name, fn
        def __call__(self, *a):
            if not on[0]:
                return self.fn(*a)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); rc = self.fn(*a); e1.record()
            rec.append((self.name, tuple(int(v) for v in a if isinstance(v, int) and abs(v) < (1 << 20)), e0, e1))
            return rc
    proxy = type("LibProxy", (), {})()
    for name in _C.SIGNATURES:
        fn = getattr(L, name)
        setattr(proxy, name, Timed(name, fn) if fn.restype is _C.c_int and name not in ("ttdg_version", "ttdg_limit") and "supported" not in name else fn)
    _C._lib = proxy
    def step():
        m.train()
        loss, _, _, _ = m(inputs, branch="TTT")
        opt.zero_grad(); loss.backward(); opt.step(1)
        m.eval()
        return m(inputs)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    on[0] = True
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, ints, e0, e1 in rec:
        a = agg[(name, ints)]
        a[0] += 1; a[1] += e0.elapsed_time(e1)
    def flops(name, a):
        if name == "ttdg_conv_tc":          # res_mode, relu, flip, N, H, W, Cin, Cout, R, S, pad, stride, out_stride, outH, outW
            _, _, _, N, H, W, Cin, Cout, R, S, pad, stride = a[:12]
            Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
            return 2.0 * N * Ho * Wo * Cin * Cout * R * S
        if name == "ttdg_wgrad_tc":         # precise, N, H, W, Cin, Cout, R, S, stride, pad
            _, N, H, W, Cin, Cout, R, S, stride, pad = a[:10]
            Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
            return 2.0 * N * Ho * Wo * Cin * Cout * R * S
        if name in ("ttdg_conv_wgrad", "ttdg_conv_dgrad"):   # N, H, W, Cin, Cout, R, S, stride, pad
            N, H, W, Cin, Cout, R, S, stride, pad = a[:9]
            Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
            return 2.0 * N * Ho * Wo * Cin * Cout * R * S
        if name == "ttdg_conv_fwd":         # res_mode, relu, N, H, W, Cin, Cout, R, S, stride, pad
            _, _, N, H, W, Cin, Cout, R, S, stride, pad = a[:11]
            Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
            return 2.0 * N * Ho * Wo * Cin * Cout * R * S
        return 0.0
    tot = sum(v[1] for v in agg.values()) / reps
    print("sum of timed library calls: %.2f ms / step (%d calls / step)" % (tot, len(rec) // reps))
    print("entry,geometry,calls_per_step,ms_per_step,us_per_call,tflops_fp32_equiv")
    for (name, a), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 70]:
        f = flops(name, a)
        print('%s,"%s",%.1f,%.3f,%.1f,%s' % (name, " ".join(map(str, a)), n / reps, ms / reps, 1e3 * ms / n, ("%.1f" % (f * n / (ms * 1e-3) / 1e12 / 1.0) if f else "")))
elif what == "conv":
    # isolated timing of conv_tc geometries (random data): python tools/run_kernels.py conv <reps> [N,H,W,Cin,Cout,R ...]
    from ttdg_b200 import detector
    geoms = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or [(8, 128, 128, 256, 256, 3), (8, 32, 32, 256, 256, 3),
                                                                             (800, 14, 14, 256, 256, 3), (8000, 1, 1, 12544, 1024, 1),
                                                                             (8, 128, 128, 64, 256, 1), (8, 32, 32, 256, 1024, 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (N, H, W, Cin, Cout, R) in geoms:
        x = torch.randn(N, H, W, Cin, device=dev)
        w = torch.randn(R, R, Cin, Cout, device=dev) * 0.05
        holder = torch.nn.Module()
        y = detector.conv_forward(x, w, None, None, None, 0, False, R, R, 1, R // 2, owner=holder)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(reps):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            detector.conv_forward(x, w, None, None, None, 0, False, R, R, 1, R // 2, out=y, owner=holder)
            e1.record(); torch.cuda.synchronize()
            ms += e0.elapsed_time(e1) / reps
        fl = 2.0 * N * H * W * Cin * Cout * R * R
        print("conv %s: %.1f us, %.1f TFLOP/s fp32-equivalent" % ((N, H, W, Cin, Cout, R), ms * 1e3, fl / (ms * 1e-3) / 1e12))
elif what == "gagm_fixed":
    # GA-GM solver alone on fixed seeded inputs (init affinity state: ~212 iterations, 200 of them Hungarian), CUDA events
    from adapteacher.modeling.GModule.multi_graph_matching import MGM3_unsup
    m = MGM3_unsup(2, 32).to(dev)
    m.load_state_dict(synth.mgm_unsup_state(0))
    m.train()
    for sizes, seed in (((33, 34, 33, 33, 46, 30, 34, 38), 77), ((30, 31, 29, 32, 28, 30, 31, 32), 5), ((40, 44, 46, 41, 39, 45, 43, 42), 9)):
        nodes, labels, _ = synth.mgm_inputs(sizes, seed)
        U = synth.universe(0).to(dev)
        with torch.no_grad():
            m([n.to(dev) for n in nodes], [l.to(dev) for l in labels], U)
        aux = m.last_aux
        from ttdg_b200 import _C
        for mode in ((3,) if os.environ.get('TTDG_FIXED_MODE3') else (0, 3, 4)):
            prev = _C.lib().ttdg_gagm_set_lap_fast(mode)
            ms = 0.0
            for _ in range(reps + 1):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                U2, info = ops.gagm_solve(aux["A"], aux["Wds"], aux["U0"], list(sizes), return_info=True)
                e1.record(); torch.cuda.synchronize()
                if _ > 0:
                    ms += e0.elapsed_time(e1) / reps
            _C.lib().ttdg_gagm_set_lap_fast(prev)
            inf = info.tolist()
            print("gagm_fixed lap_fast %d sizes %s: %.3f ms, iterations %d (sinkhorn %d, hungarian %d), %.1f us / hungarian iteration incl. everything, graph-0 LAP steps %d, fast-path fall-backs %d; CTA-0 kcycles: kernel %d, hungarian stage %d, in LAP %d, barrier wait %d; sinkhorn stage: phase 1 %d, V %d, projector %d"
                  % (mode, sizes, ms, inf[0], inf[1], inf[2], 1e3 * ms / max(inf[2], 1), inf[5], inf[7], inf[8], inf[9], inf[10], inf[11], inf[12], inf[13], inf[14]))
            if mode == 3:
                import ctypes
                seg = (ctypes.c_int64 * 24)()
                _C.lib().ttdg_gagm_read_profile(seg)
                print("gagm_fixed lap_fast 3   hungarian-stage kcycles of CTA 0: T build %d, Q gather %d, V1 %d, V2 %d, V store %d, projection %d, U store + norms %d, "
                      "norm reduce %d, barrier %d, tail %d; LAP: auction scans %d, bids %d, augmentations %d, certificate %d (reach %d, Kahn %d); free rows per round %s"
                      % (tuple(int(v) >> 10 for v in seg[:16]) + (str([int(v) for v in seg[16:21]]),)))
elif what == "gagm_bench":
    # the GA-GM solver on the BENCH workload's own problems: run the bench step a few times (the weights drift), then time the
    # solver alone on the step's (A, Wds, U0) under every LAP mode
    sys.path.insert(0, ROOT)
    import bench
    from ttdg_b200 import _C
    m, opt = bench.build_ours(dev)
    inputs = [dict(d, image=d["image"].to(dev)) for d in bench.make_inputs(0)]
    for step_i in range(16):
        m.train()
        loss, _, _, _ = m(inputs, branch="TTT")
        opt.zero_grad(); loss.backward(); opt.step(1)
        if step_i in (3, 9, 15):
            aux = m.multi_matching_unsup.last_aux
            sizes = list(aux["sizes"])
            for mode in (3,):
                prev = _C.lib().ttdg_gagm_set_lap_fast(mode)
                ms = 0.0
                for r in range(reps + 1):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    U2, info = ops.gagm_solve(aux["A"], aux["Wds"], aux["U0"], sizes, return_info=True)
                    e1.record(); torch.cuda.synchronize()
                    if r > 0:
                        ms += e0.elapsed_time(e1) / reps
                _C.lib().ttdg_gagm_set_lap_fast(prev)
                inf = info.tolist()
                print("gagm_bench step %d lap_fast %d sizes %s: %.3f ms, iterations %d (sinkhorn %d, hungarian %d), graph-0 LAP steps %d, fall-backs %d; "
                      "CTA-0 kcycles: kernel %d, hungarian stage %d, in LAP %d, barrier wait %d; sinkhorn stage: phase 1 %d, V %d, projector %d" %
                      (step_i, mode, sizes, ms, inf[0], inf[1], inf[2], inf[5], inf[7], inf[8], inf[9], inf[10], inf[11], inf[12], inf[13], inf[14]))
                import ctypes
                seg = (ctypes.c_int64 * 24)()
                _C.lib().ttdg_gagm_read_profile(seg)
                print("gagm_bench   hungarian-stage kcycles of CTA 0: T build %d, Q gather %d, V1 %d, V2 %d, V store %d, projection %d, U store + norms %d, "
                      "norm reduce %d, barrier %d, tail %d; LAP: auction scans %d, bids %d, augmentations %d, certificate %d (reach %d, Kahn %d); free rows per round %s" % (tuple(int(v) >> 10 for v in seg[:16]) + (str([int(v) for v in seg[16:21]]),)))
elif what == "busy":
    # hot (not cold-cache) per-kernel device time of the full step and the GPU-busy fraction, from CUPTI via torch.profiler
    sys.path.insert(0, ROOT)
    import time, collections
    import bench
    from torch.profiler import profile, ProfilerActivity
    m, opt = bench.build_ours(dev)
    host = "host" in sys.argv[3:]                          # e2e form: pinned host images, result read back every step
    inputs = [dict(d, image=(d["image"].pin_memory() if host else d["image"].to(dev))) for d in bench.make_inputs(0)]
    host_res = torch.empty(9, dtype=torch.float64).pin_memory()
    def step():
        m.train()
        loss, _, _, _ = m(inputs, branch="TTT")
        opt.zero_grad(); loss.backward(); opt.step(1)
        m.eval()
        out = m(inputs)
        if host:
            res = torch.stack([o["instances"].pred_masks.sum().to(torch.float64) for o in out] + [loss.detach().to(torch.float64)])
            host_res.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out
    if "overlap" in sys.argv[3:]:
        # the product's schedule (adapteacher/engine/trainer.py OverlappedEval, as bench.py runs it): pass 2 of the previous
        # step inside this step's solver window, on a second stream, from a weight snapshot
        from adapteacher.engine.trainer import OverlappedEval
        pipe = OverlappedEval(m)

        class Sink:
            def reset(self): self.out = []
            def process(self, i, o): self.out += o
            def evaluate(self): self.out = []

        sink = Sink()

        def step():
            m.train()
            ops.SOLVER_WINDOW_HOOK[0] = pipe.window if pipe.active else None
            loss, _, _, _ = m(inputs, branch="TTT")
            ops.SOLVER_WINDOW_HOOK[0] = None
            opt.zero_grad(); loss.backward(); opt.step(1)
            if pipe.active:
                pipe.drain()
            pipe.begin("busy", [inputs], sink)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for e in prof.events():
        if str(e.device_type).endswith("CUDA"):
            name = e.name.split("(")[0][:80]
            tot[name] += e.device_time / reps / 1e3
            cnt[name] += 1
    busy = sum(tot.values())
    print("wall ms/step %.2f (under the profiler), device busy ms/step %.2f (%.0f%%), launches/step %d" %
          (wall, busy, 100 * busy / wall, sum(cnt.values()) // reps))
    if "gaps" in sys.argv[3:]:
        # idle time of the device between consecutive kernels, attributed to the kernel that FOLLOWS the gap
        evs = sorted(((e.time_range.start, e.time_range.end, e.name.split("(")[0][:70]) for e in prof.events()
                      if str(e.device_type).endswith("CUDA")), key=lambda t: t[0])
        gap_after, gap_n = collections.defaultdict(float), collections.Counter()
        end = evs[0][1]
        for (st, en, name), prev in zip(evs[1:], evs[:-1]):
            g = st - end
            if g > 2.0:
                key = "%s  <-after-  %s" % (name, prev[2])
                gap_after[key] += g / reps / 1e3; gap_n[key] += 1
            end = max(end, en)
        print("idle ms/step %.2f in %d gaps > 2 us" % (sum(gap_after.values()), sum(gap_n.values()) // reps))
        print("gap_ms_per_step,gaps_per_step,kernel_after_gap <-after- kernel_before")
        for k, v in sorted(gap_after.items(), key=lambda x: -x[1])[:40]:
            print("%.3f,%.1f,%s" % (v, gap_n[k] / reps, k))
    print("kernel,launches_per_step,ms_per_step,share_pct")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:45]:
        print('"%s",%d,%.3f,%.1f' % (k, cnt[k] // reps, v, 100 * v / busy))
torch.cuda.synchronize()
