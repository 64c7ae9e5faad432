"""Generate tests/golden/*.npz from the reference's OWN Python (imported via ref_shim).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python -m oracle.gen_golden            # writes tests/golden/mgm_<case>_<variant>.npz etc.

Inputs are NOT stored: they are re-drawn from ``ttdg_b200.synth`` seeds; each file carries an input
checksum so RNG drift is detected.  Records the library versions the vectors were made with
(SciPy 1.18.1 here; the reference pins 1.7.3 - SURVEY Appendix C).
"""
import os
import sys

import numpy as np
import scipy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load_synth():
    # import synth.py by path so our 'adapteacher' mirror never enters this process
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ttdg_synth", os.path.join(ROOT, "ttdg-mgm_b200", "ttdg_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def checksum(tensors):
    s = 0.0
    for t in tensors:
        t = t.double()
        s += float(t.sum()) + float((t * t).sum())
    return s


def summarise(name, t):
    """Big parameter gradients are stored as a strided sub-sample plus float64 moments."""
    t = t.detach()
    if t.numel() <= 4096:
        return {name: t.numpy()}
    sub = t[::16, ::16] if t.dim() == 2 else t[::16]
    return {name + "_sub16": sub.contiguous().numpy(), name + "_sum": np.array(float(t.double().sum())),
            name + "_sumsq": np.array(float((t.double() ** 2).sum()))}


def run_mgm_case(ref, synth, sizes, seed, variant, univ_seed=0):
    torch.manual_seed(0)
    model = ref.MGM3_unsup(2, 32)
    sd = synth.mgm_unsup_state(0)
    if variant == "pert":
        sd = synth.perturb_affinity_state(sd, 0)
    missing = model.load_state_dict(sd, strict=True)
    model.train()                                     # TTT runs in train mode
    nodes, labels, masks = synth.mgm_inputs(sizes, seed)
    U = synth.universe(univ_seed)
    from oracle.ref_shim import MaskedDropout
    model.intra_domain_graph.dot_product_attention.dropout = MaskedDropout(masks)
    # the output-projection dropout (attentions.py:57,85) feeds a discarded tensor: make it a no-op
    model.intra_domain_graph.dropout = torch.nn.Identity()

    cap = {}
    ga = model.ga_mgmc
    orig_fwd = ga.forward

    def fwd(A, W, U0, ms, n_univ, quad_weight=1., cluster_quad_weight=1., num_clusters=1):
        cap["A"], cap["W"], cap["U0"] = A.detach().clone(), W.detach().clone(), U0.detach().clone()
        out = orig_fwd(A, W, U0, ms, n_univ, quad_weight, cluster_quad_weight, num_clusters)
        cap["U"] = out[0].detach().clone()
        return out
    ga.forward = fwd
    counts = {"hung": 0, "sk": 0}
    trace = []                       # one record per GA-GM iteration: U_t, projector, tau
    in_gagm = {"on": False}
    orig_h = ref.mgm.hungarian

    def hung(s, *a, **k):
        counts["hung"] += 1
        if in_gagm["on"] and trace:
            trace[-1]["proj"] = 1
        return orig_h(s, *a, **k)
    ref.mgm.hungarian = hung
    orig_chain = torch.chain_matmul

    def chain(*a):
        counts["sk"] += 1            # one chain_matmul per GA-GM iteration (HiPPI unused here)
        trace.append({"U": a[3].detach().clone(), "proj": -1, "tau": 0.0})
        return orig_chain(*a)
    torch.chain_matmul = chain
    orig_sk = ref.mgm.Sinkhorn

    class SkSpy(orig_sk):            # gagm builds a fresh Sinkhorn(tau=...) per sinkhorn-projected iteration (mgm:333-349)
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            if in_gagm["on"] and trace:
                trace[-1]["proj"], trace[-1]["tau"] = 0, float(k.get("tau"))
    ref.mgm.Sinkhorn = SkSpy
    orig_gagm = ga.gagm

    def gagm_spy(*a, **k):
        in_gagm["on"] = True
        try:
            return orig_gagm(*a, **k)
        finally:
            in_gagm["on"] = False
    ga.gagm = gagm_spy
    try:
        nodes_g = [n.clone().requires_grad_(True) for n in nodes]
        loss = model(nodes_g, labels, U)
        loss.backward()
    finally:
        ref.mgm.hungarian = orig_h
        torch.chain_matmul = orig_chain
        ref.mgm.Sinkhorn = orig_sk
    out = {
        "sizes": np.array(sizes), "seed": np.array(seed), "variant": np.array(variant),
        "input_checksum": np.array(checksum(nodes + [U] + masks)),
        "loss": np.array(loss.detach().numpy()),
        "A": cap["A"].numpy(), "Wds": cap["W"].numpy(), "U0": cap["U0"].numpy(),
        "U": cap["U"].numpy().astype(np.uint8),
        "U_is_binary": np.array(bool(((cap["U"] == 0) | (cap["U"] == 1)).all())),
        "gagm_iters": np.array(counts["sk"]), "hungarian_calls": np.array(counts["hung"]),
        "scipy_version": np.array(scipy.__version__), "torch_version": np.array(torch.__version__),
    }
    for i, n in enumerate(nodes_g):
        out[f"grad_nodes_{i}"] = n.grad.numpy()
    for k, p in model.node_affinity.named_parameters():
        out.update(summarise("grad_aff_" + k, p.grad))
    if sum(sizes) > 300:             # keep the fixture small
        del out["A"], out["Wds"], out["U0"]
    elif variant == "pert":
        # teacher-forcing samples of the reference's own fp32 trajectory: (U_t, projector, tau) -> U_{t+1}
        # (the full trajectory is chaotic - see mgm_port.gagm - single steps are not)
        T = len(trace)
        sk_idx = [t for t in range(T) if trace[t]["proj"] == 0]
        hg_idx = [t for t in range(T) if trace[t]["proj"] == 1]
        pick = sorted(set(sk_idx[:2] + sk_idx[len(sk_idx) // 2:len(sk_idx) // 2 + 1] + sk_idx[-2:] +
                          hg_idx[:2] + hg_idx[len(hg_idx) // 2:len(hg_idx) // 2 + 1] + hg_idx[-1:]))
        nxt = lambda t: trace[t + 1]["U"] if t + 1 < T else cap["U"]
        out["trace_iter"] = np.array(pick)
        out["trace_proj"] = np.array([trace[t]["proj"] for t in pick])
        out["trace_tau"] = np.array([trace[t]["tau"] for t in pick])
        for k, t in enumerate(pick):
            out[f"trace_Uin_{k}"] = trace[t]["U"].numpy()
            out[f"trace_Uout_{k}"] = nxt(t).numpy()
        out["trace_proj_all"] = np.array([r["proj"] for r in trace], dtype=np.int8)
    return out


def run_op_goldens(ref, synth):
    """Small per-operator vectors from the reference modules (Sinkhorn wrapper, hungarian, Affinity,
    focal-BCE permutation loss, attention adjacency, node sampler)."""
    g = torch.Generator().manual_seed(4242)
    out = {}
    # Sinkhorn: per-item differentiable path with dummy rows (mgm:467-468,519-522)
    sk = ref.Sinkhorn(max_iter=20, tau=0.05, epsilon=1e-10, batched_operation=False)
    for name, (n1, n2) in {"sk_23x40": (23, 40), "sk_32x32": (32, 32), "sk_5x7": (5, 7)}.items():
        s = torch.randn(n1, n2, generator=g).requires_grad_(True)
        y = sk(s, dummy_row=True)
        w = torch.randn(n1, n2, generator=g)
        (y * w).sum().backward()
        out[name + "_in"], out[name + "_out"] = s.detach().numpy(), y.detach().numpy()
        out[name + "_w"], out[name + "_grad"] = w.numpy(), s.grad.numpy()
    # batched, no-grad GA-GM projector calls (mgm:330-353)
    skb = ref.Sinkhorn(max_iter=20, tau=0.1, batched_operation=True)
    v = torch.randn(4, 20, 32, generator=g)
    out["skb_eq_le_in"], out["skb_eq_le_out"] = v.numpy(), skb(v, dummy_row=True).numpy()
    v = torch.randn(3, 45, 32, generator=g)
    out["skb_eq_gt_in"] = v.numpy()
    out["skb_eq_gt_out"] = skb(v.transpose(1, 2), dummy_row=True).transpose(1, 2).numpy()
    sizes = [32, 45, 20]
    vs = [torch.randn(n, 32, generator=g) for n in sizes]
    padded = torch.stack(ref.pad_tensor.pad_tensor(vs), dim=0)
    out["skb_rag_in"], out["skb_rag_sizes"] = padded.numpy(), np.array(sizes)
    out["skb_rag_out"] = skb(padded, torch.tensor(sizes), dummy_row=True).numpy()
    sizes2 = [17, 32, 20]
    vs = [torch.randn(n, 32, generator=g) for n in sizes2]
    padded = torch.stack(ref.pad_tensor.pad_tensor(vs), dim=0)
    out["skb_rag2_in"], out["skb_rag2_sizes"] = padded.numpy(), np.array(sizes2)
    out["skb_rag2_out"] = skb(padded, torch.tensor(sizes2), dummy_row=True).numpy()
    # plain square / wide, no dummy row (microbench form; U_sup's sinkhorn(U) at mgm:143 is tall)
    sk50 = ref.Sinkhorn(max_iter=50, tau=0.05)
    s = torch.randn(2, 48, 48, generator=g)
    out["sk50_in"], out["sk50_out"] = s.numpy(), sk50(s).numpy()
    s = torch.randn(70, 32, generator=g)
    out["sk_tall_in"], out["sk_tall_out"] = s.numpy(), sk(s).numpy()
    # hungarian (utils/hungarian.py:8-65)
    for name, shp in {"hung_40x32": (40, 32), "hung_20x32": (20, 32), "hung_32x32": (32, 32),
                      "hung_23x57": (23, 57)}.items():
        s = torch.randn(*shp, generator=g)
        out[name + "_in"], out[name + "_out"] = s.numpy(), ref.hungarian_fn(s).numpy().astype(np.uint8)
    s = torch.randint(0, 3, (12, 9), generator=g).float()          # tie-heavy
    out["hung_ties_in"], out["hung_ties_out"] = s.numpy(), ref.hungarian_fn(s).numpy().astype(np.uint8)
    # Affinity (utils/affinity.py:44-57) with the synthetic state
    aff = ref.Affinity(256)
    sd = synth.perturb_affinity_state(synth.mgm_unsup_state(0), 0)
    aff.load_state_dict({k[len("node_affinity."):]: v for k, v in sd.items() if k.startswith("node_affinity.")})
    X = torch.randn(37, 256, generator=g).requires_grad_(True)
    Y = torch.randn(52, 256, generator=g).requires_grad_(True)
    M = aff(X, Y)
    wM = torch.randn(37, 52, generator=g)
    (M * wM).sum().backward()
    out["aff_X"], out["aff_Y"], out["aff_M"], out["aff_w"] = X.detach().numpy(), Y.detach().numpy(), M.detach().numpy(), wM.numpy()
    out["aff_dX"], out["aff_dY"] = X.grad.numpy(), Y.grad.numpy()
    for k, p in aff.named_parameters():
        out.update(summarise("aff_d_" + k, p.grad))
    # PermutationLoss -> BCEFocalLoss (utils/losses.py:83-103,419-455)
    crit = ref.PermutationLoss()
    S = torch.rand(23, 40, generator=g).requires_grad_(True)
    Yp = (torch.rand(23, 40, generator=g) > 0.9).float()
    l = crit(S, Yp, torch.tensor(23), torch.tensor(40))
    l.backward()
    out["focal_S"], out["focal_Y"], out["focal_loss"], out["focal_dS"] = S.detach().numpy(), Yp.numpy(), l.detach().numpy(), S.grad.numpy()
    # attention adjacency (utils/attentions.py:60-86) with injected dropout mask
    from oracle.ref_shim import MaskedDropout
    att = ref.MultiHeadAttention(256, 1, dropout=0.1, version="v2")
    att.load_state_dict({k[len("intra_domain_graph."):]: v for k, v in sd.items() if k.startswith("intra_domain_graph.")})
    att.train()
    x = torch.randn(29, 256, generator=g)
    mask = (torch.rand(29, 29, generator=g) >= 0.1).float()
    att.dot_product_attention.dropout = MaskedDropout([mask])
    att.dropout = torch.nn.Identity()
    _, adj = att([x, x, x])
    out["att_x"], out["att_mask"], out["att_adj"] = x.numpy(), mask.numpy(), adj.detach().numpy()
    att.eval()
    att.dot_product_attention.dropout = torch.nn.Dropout(0.1)
    att.dot_product_attention.dropout.eval()
    _, adj = att([x, x, x])
    out["att_adj_eval"] = adj.detach().numpy()
    return out


class FakeBoxes:
    def __init__(self, t):
        self.tensor = t


class FakeInstances:
    """Duck-typed Detectron2 ``Instances`` (build_graph.py:78-85 only touches these members)."""

    def __init__(self, boxes, classes):
        self.pred_boxes = FakeBoxes(boxes)
        self.pred_classes = classes
        self._fields = {"pred_boxes": self.pred_boxes, "pred_classes": classes}

    def __len__(self):
        return self.pred_boxes.tensor.shape[0]


def run_sampler_goldens(ref, synth):
    """PrototypeComputation (build_graph.py:160-250) on small seeded FPN pyramids (features are
    re-drawn from ``synth.sampler_case``; only outputs are stored)."""
    out = {}
    pc = ref.PrototypeComputation(2, 10)
    for name in synth.SAMPLER_CASES:
        c = synth.SAMPLER_CASES[name]
        S, B = c["S"], len(c["boxes"])
        feats = synth.sampler_feats(name)
        targets = [FakeInstances(torch.tensor(b, dtype=torch.float32).reshape(-1, 4),
                                 torch.tensor(k, dtype=torch.int64)) for b, k in zip(c["boxes"], c["classes"])]
        nodes, labels = pc(feats, targets)
        out[name + "_input_checksum"] = np.array(checksum(feats))
        out[name + "_nout"] = np.array(len(nodes))
        for i, (n, l) in enumerate(zip(nodes, labels)):
            out[f"{name}_nodes{i}"] = n.numpy()
            out[f"{name}_labels{i}"] = l.numpy()
    return out


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    ref = ref_shim.load()
    synth = _load_synth()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)     # fixed reduction order in the generating run
    which = sys.argv[1:] or ["mgm", "ops", "sampler"]
    if "mgm" in which:
        for name, (sizes, seed) in synth.MGM_CASES.items():
            for variant in ("init", "pert"):
                if variant == "init" and name not in ("g4x40", "ragged4", "g2x32"):
                    continue
                res = run_mgm_case(ref, synth, sizes, seed, variant)
                # keep files small: drop A/Wds for the big case
                np.savez_compressed(os.path.join(GOLDEN, f"mgm_{name}_{variant}.npz"), **res)
                print(name, variant, "loss", float(res["loss"]), "iters", int(res["gagm_iters"]),
                      "hung", int(res["hungarian_calls"]), "ones", int(res["U"].sum()), flush=True)
    if "ops" in which:
        np.savez_compressed(os.path.join(GOLDEN, "ops.npz"), **run_op_goldens(ref, synth))
        print("ops done")
    if "sampler" in which:
        np.savez_compressed(os.path.join(GOLDEN, "sampler.npz"), **run_sampler_goldens(ref, synth))
        print("sampler done")


if __name__ == "__main__":
    main()
