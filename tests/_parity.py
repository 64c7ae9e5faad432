"""Parity of the ASSEMBLED hot path on BASELINE.json's own shapes, CUDA path vs oracle/ttt_port (checker).

The path has two kinds of stages:
  * continuous ones (convolutions, RoIAlign, FC, affinity, Sinkhorn, loss, every gradient): compared numerically;
  * discrete ones (top-k, NMS, score threshold, node sampling, the GA-GM solver's 200 Hungarian projections): a 1e-6
    difference upstream can flip a decision, after which two runs legitimately differ.  The reference has the same
    property against itself (fp32 vs float64, or 1 vs 8 MKL threads: DESIGN section 3).
So every quantity is measured twice: free-running (each side follows its own decisions; reported together with how much
the fp32 restatement itself moves against its float64 limit) and teacher-forced (the discrete result of one side is fed to
the other, which leaves only continuous arithmetic: tight tolerances).

Used by tests/test_gpu_parity_configs.py, __graft_entry__.smoke() and bench.py's cpu_baseline leg (`parity` key)."""
import time

import numpy as np
import torch

from oracle import detector_port as dp
from oracle import ttt_port
from ttdg_b200 import synth

CONFIGS = {
    # BASELINE.json configs[1]: batch = 8 synthetic 512 x 512 fundus-like 2-class images, Sinkhorn 20 iters, fp32
    "configs1": dict(size=512, batch=8, num_classes=2, polyp=False, first=0),
    # BASELINE.json configs[3]: 384 x 384 polyp-like 1-class, 5-graph matching problem (TEST.BATCH 5)
    "configs3": dict(size=384, batch=5, num_classes=1, polyp=True, first=500),
    # small shape for smoke()
    "smoke": dict(size=128, batch=3, num_classes=2, polyp=False, first=200),
}


def keep_masks_fn(seed=7, p=0.1):
    def fn(sizes):
        g = torch.Generator().manual_seed(seed)
        return [(torch.rand(n, n, generator=g) >= p).to(torch.float32) for n in sizes]
    return fn


def iou(a, b):
    inter = (a & b).flatten(1).sum(1).double()
    union = (a | b).flatten(1).sum(1).double()
    return torch.where(union > 0, inter / union, torch.ones_like(union))


def miou_vs_gt(pred, gt):
    """mean over predicted instances of the best IoU against the ground-truth masks (the harness computes
    mIoU itself from the same masks, over all detections)."""
    if len(pred) == 0:
        return float("nan")
    best = torch.stack([iou(pred, gt[j:j + 1].expand_as(pred)) for j in range(len(gt))]).max(0)[0]
    return float(best.mean())


def match(a, b, tol):
    """For every box of a: index of the nearest box of b (L-inf) and whether it is within tol."""
    if len(a) == 0 or len(b) == 0:
        return torch.zeros(len(a), dtype=torch.long), torch.zeros(len(a), dtype=torch.bool)
    d = (a[:, None, :].double() - b[None, :, :].double()).abs().max(-1)[0]
    mn, idx = d.min(1)
    return idx, mn < tol


def det_match(da, db, tol=0.1):
    """fraction of detections of `da` (boxes, scores, classes) with a partner in `db`: same class, box within tol px."""
    idx, ok = match(da[0].cpu(), db[0].cpu(), tol)
    ok = ok & (da[2].cpu() == db[2].cpu()[idx]) if len(db[0]) else ok
    return idx, ok


def rel_l2(a, r):
    a, r = a.double(), r.double()
    n = float(r.norm())
    return float((a - r).norm()) / n if n > 0 else float((a - r).norm())


def build_model(num_classes, state="bench"):
    from adapteacher.modeling.meta_arch.rcnn import DAobjTwoStagePseudoLabGeneralizedRCNN
    m = DAobjTwoStagePseudoLabGeneralizedRCNN(num_classes).cuda()
    sd_det = synth.detector_state_calibrated(0, num_classes)
    sd_mgm = synth.mgm_unsup_state(0)               # the reference constructors' init: what bench.py loads
    U = synth.universe(0)
    sd = dict(sd_det)
    sd.update({"multi_matching_unsup." + k: v for k, v in sd_mgm.items()})
    sd["multi_matching_sup.U"] = U
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("D_img.") for k in missing), (missing, unexpected)
    return m, sd_det, sd_mgm, U


def gpu_named_grads(m):
    """{d2 / reference state-dict key: gradient in the reference's layout} for the adapted parameters of the CUDA model."""
    out = {}
    mods = dict(m.named_modules())
    for name, p in m.named_parameters():
        if p.grad is None:
            continue
        mod_name, kind = name.rsplit(".", 1)
        layer = mods[mod_name]
        g = p.grad.detach()
        if hasattr(layer, "cout_p"):                       # detector.Conv2d: [R][S][Cin_p][Cout_p] -> Cout x Cin x R x S
            g = g[:, :, :layer.cin, :layer.cout].permute(3, 2, 0, 1) if kind == "weight" else g[:layer.cout]
        out[name] = g.cpu().contiguous()
    return out


def images_of(cfg):
    ims = [synth.fundus_like_image(cfg["first"] + i, cfg["size"], cfg["polyp"]) for i in range(cfg["batch"])]
    return ims, [im["image"] for im in ims]


@torch.no_grad()
def eval_parity(m, sd_det, cfg, with_f64=True, log=print):
    """Eval pass (GeneralizedRCNN.inference) on the SAME weights on both sides."""
    ims, images = images_of(cfg)
    S = cfg["size"]
    rep = {}
    t0 = time.time()
    m.eval()
    det = m._det[0]
    res_g, feats_g, props_g, dets_g = det.inference(images)
    torch.cuda.synchronize()
    res_o, feats_o, props_o, dets_o, probs_o = dp.inference(sd_det, images)
    log("  eval: oracle fp32 %.1f s" % (time.time() - t0))
    gts = [im["gt_masks"] for im in ims]

    def free_running(ra, rb):
        mf, d_all, d_m, ious = [], [], [], []
        for n in range(len(images)):
            a, b = ra[n], rb[n]
            idx, ok = det_match((a["pred_boxes"], a["scores"], a["pred_classes"]), (b["pred_boxes"], b["scores"], b["pred_classes"]))
            mf.append(float(ok.float().mean()) if len(ok) else 1.0)
            pa, pb = a["pred_masks"].cpu(), b["pred_masks"].cpu()
            d_all.append(miou_vs_gt(pa, gts[n]) - miou_vs_gt(pb, gts[n]))
            if ok.any():
                d_m.append(miou_vs_gt(pa[ok], gts[n]) - miou_vs_gt(pb[idx][ok], gts[n]))
                ious.append(iou(pa[ok], pb[idx][ok]))
        ious = torch.cat(ious) if ious else torch.ones(1, dtype=torch.float64)
        return {"matched_frac": float(np.mean(mf)), "miou_delta_all": float(np.mean(np.abs(d_all))),
                "miou_delta_matched": float(np.mean(np.abs(d_m))) if d_m else float("nan"),
                "mask_iou_matched_mean": float(ious.mean()), "mask_iou_matched_min": float(ious.min())}

    rep["free_running_gpu_vs_fp32"] = free_running(res_g, res_o)
    # ---- teacher-forced, stage by stage (continuous arithmetic only)
    # backbone + FPN
    def pyr(fa, fb):
        return max(float((a.double() - b.double()).abs().max() / b.double().abs().max()) for a, b in zip(fa, fb))
    feats_g_nchw = [f.permute(0, 3, 1, 2).cpu() for f in feats_g]
    rep["pyramid_rel_max"] = pyr(feats_g_nchw, feats_o)
    # box head on the ORACLE's proposals
    props_forced = [(b.cuda(), s.cuda()) for b, s in props_o]
    dets_f = m.roi_heads.forward_box(feats_g, props_forced, (S, S))
    mf, sd_, bd = [], [], []
    for n in range(len(images)):
        idx, ok = det_match(dets_f[n], dets_o[n])
        mf.append(float(ok.float().mean()))
        sd_.append(float((dets_f[n][1].cpu()[ok] - dets_o[n][1][idx][ok]).abs().max()) if ok.any() else 0.0)
        bd.append(float((dets_f[n][0].cpu()[ok] - dets_o[n][0][idx][ok]).abs().max()) if ok.any() else 0.0)
    rep["box_head_forced_proposals"] = {"matched_frac": float(np.mean(mf)), "score_max_abs": max(sd_), "box_max_abs_px": max(bd)}
    # mask branch + paste on the ORACLE's detections
    dets_forced = [(b.cuda(), s.cuda(), c.cuda()) for b, s, c in dets_o]
    res_f = m.roi_heads.forward_mask(feats_g, dets_forced, (S, S), (S, S))
    d, ious, flips, px = [], [], 0, 0
    for n in range(len(images)):
        pa, pb = res_f[n]["pred_masks"].cpu(), res_o[n]["pred_masks"]
        assert pa.shape == pb.shape, (pa.shape, pb.shape)
        d.append(miou_vs_gt(pa, gts[n]) - miou_vs_gt(pb, gts[n]))
        ious.append(iou(pa, pb))
        flips += int((pa != pb).sum())
        px += int((pa | pb).sum())
    ious = torch.cat(ious)
    rep["mask_branch_forced_detections"] = {"miou_delta": float(np.mean(np.abs(d))), "miou_delta_max_image": float(np.max(np.abs(d))),
                                            "mask_iou_mean": float(ious.mean()), "mask_iou_min": float(ious.min()),
                                            "flipped_pixels": flips, "mask_pixels": px, "instances": int(len(ious))}
    if with_f64:
        t1 = time.time()
        with dp.float64():
            res_64, feats_64 = dp.inference({k: v.double() for k, v in sd_det.items()}, images)[:2]
        log("  eval: oracle float64 %.1f s" % (time.time() - t1))
        rep["pyramid_rel_max_gpu_vs_f64"] = pyr(feats_g_nchw, feats_64)
        rep["pyramid_rel_max_fp32_vs_f64"] = pyr(feats_o, feats_64)
        rep["free_running_fp32_vs_f64"] = free_running(res_o, res_64)
        rep["free_running_gpu_vs_f64"] = free_running(res_g, res_64)
    return rep


def ttt_parity(m, sd_det, sd_mgm, U, cfg, with_f64=True, log=print):
    """One adaptation step: loss, gradients of every adapted tensor and the post-step weights."""
    from ttdg_b200.optim import FlatSGD
    ims, images = images_of(cfg)
    rep = {}
    opt = FlatSGD(m.adapted_parameters(), lr=0.005, momentum=0.9, weight_decay=1e-4)
    p0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    m.train()
    m.multi_matching_unsup.debug_keep_masks = keep_masks_fn()
    inputs = [{"image": im, "height": cfg["size"], "width": cfg["size"], "image_id": i} for i, im in enumerate(images)]
    loss_g, _, _, _ = m(inputs, branch="TTT")
    assert loss_g is not None and torch.isfinite(loss_g)
    opt.zero_grad()
    loss_g.backward()
    torch.cuda.synchronize()
    aux = m.multi_matching_unsup.last_aux
    sizes = list(aux["sizes"])
    dets_g = [(b.detach().cpu(), s.detach().cpu(), c.detach().cpu()) for b, s, c in m.last_ttt["detections"]]
    U_g = aux["U"].detach().cpu()
    grads_g = gpu_named_grads(m)
    opt.step()
    torch.cuda.synchronize()
    p1 = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    m.multi_matching_unsup.debug_keep_masks = None
    rep["sizes"] = sizes
    rep["loss_gpu"] = float(loss_g)
    rep["gagm_iterations"] = int(aux["info"][0])

    def oracle(dtype):
        t0 = time.time()
        tr = ttt_port.Trainer(sd_det, sd_mgm, U, dtype=dtype)
        # free-running detections of this side (no adaptation yet), then the step with the CUDA path's discrete results forced
        loss = tr.ttt_step(images, keep_masks=keep_masks_fn(), dets_override=dets_g, U_override=U_g)
        log("  ttt: oracle %s %.1f s" % (str(dtype).split(".")[-1], time.time() - t0))
        return tr, loss

    tr32, loss32 = oracle(torch.float32)
    assert [int(n.shape[0]) for n in tr32.last["nodes"]] == sizes, "node sampler picked different locations"
    mf = [float(det_match(dets_g[n], tr32.last["dets"][n])[1].float().mean()) for n in range(len(images))]
    rep["ttt_detections_matched_frac_gpu_vs_fp32"] = float(np.mean(mf))
    rep["loss_fp32_forced"] = loss32
    rep["loss_rel_gpu_vs_fp32"] = abs(float(loss_g) - loss32) / abs(loss32)
    # intermediate tensors of the matching stage (continuous, forced detections)
    rep["A_max_abs_gpu_vs_fp32"] = float((aux["A"].cpu() - tr32.last["aux"]["A"]).abs().max())
    rep["Wds_max_abs_gpu_vs_fp32"] = float((aux["Wds"].cpu() - tr32.last["aux"]["Wds"]).abs().max())
    rep["U0_rel_max_gpu_vs_fp32"] = float((aux["U0"].cpu() - tr32.last["aux"]["U0"]).abs().max() / tr32.last["aux"]["U0"].abs().max())
    ref, ref_name = tr32, "fp32"
    if with_f64:
        tr64, loss64 = oracle(torch.float64)
        rep["loss_f64_forced"] = loss64
        rep["loss_rel_gpu_vs_f64"] = abs(float(loss_g) - loss64) / abs(loss64)
        rep["loss_rel_fp32_vs_f64"] = abs(loss32 - loss64) / abs(loss64)
        ref, ref_name = tr64, "f64"
    # gradients and post-step weights, per tensor
    rows = []
    for (k, p) in ref.named_adapted():
        gr = p.grad.detach()
        g_gpu = grads_g[k]
        e_gpu = rel_l2(g_gpu, gr)
        row = {"key": k, "gpu": e_gpu, "norm": float(gr.double().norm())}
        if with_f64:
            g32 = dict(tr32.named_adapted())[k].grad.detach()
            row["fp32"] = rel_l2(g32, gr)
        # weights after the SGD step: the UPDATE is what the step computes (the weights themselves agree trivially)
        upd_ref = (p.detach().double() - p0[k].double())
        upd_gpu = (p1[k].double() - p0[k].double())
        row["update_gpu"] = rel_l2(upd_gpu, upd_ref)
        row["weight_max_abs_gpu"] = float((p1[k].double() - p.detach().double()).abs().max())
        rows.append(row)
    rep["per_tensor"] = rows
    tot_ref = np.sqrt(sum(r["norm"] ** 2 for r in rows))
    all_rows, rows = rows, [r for r in rows if r["norm"] > 1e-9 * tot_ref]     # e.g. fc_M.2.bias: Sinkhorn is shift invariant
    e = np.array([r["gpu"] for r in rows])
    rep["grad_rel_l2_gpu_vs_%s" % ref_name] = {"max": float(e.max()), "median": float(np.median(e)), "tensors": len(rows)}
    rep["grad_bucket_rel_l2_gpu"] = float(np.sqrt(sum((r["gpu"] * r["norm"]) ** 2 for r in rows)) / tot_ref)
    u = np.array([r["update_gpu"] for r in rows])
    rep["update_rel_l2_gpu"] = {"max": float(u.max()), "median": float(np.median(u))}
    rep["weight_max_abs_gpu"] = float(max(r["weight_max_abs_gpu"] for r in all_rows))
    if with_f64:
        e32 = np.array([r["fp32"] for r in rows])
        rep["grad_rel_l2_fp32_vs_f64"] = {"max": float(e32.max()), "median": float(np.median(e32))}
        rep["grad_bucket_rel_l2_fp32"] = float(np.sqrt(sum((r["fp32"] * r["norm"]) ** 2 for r in rows)) / tot_ref)
    return rep
