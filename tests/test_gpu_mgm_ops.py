"""GPU parity of the matching-stage operators: CUDA (through the C-ABI library, via ttdg_b200.ops / the
adapteacher mirror) against the golden vectors made from the reference's own modules and against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clib, mgm_port  # noqa: E402  (checker only)
from ttdg_b200 import ops, synth  # noqa: E402

T = torch.from_numpy


def cu(x):
    return (T(x) if isinstance(x, np.ndarray) else x).cuda()


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(f"{golden_dir}/ops.npz")


@pytest.fixture(scope="module")
def sd():
    return synth.perturb_affinity_state(synth.mgm_unsup_state(0), 0)


@pytest.mark.parametrize("name", ["sk_23x40", "sk_32x32", "sk_5x7"])
def test_sinkhorn_small_fwd_bwd_golden(g, name):
    s = cu(g[name + "_in"]).requires_grad_(True)
    y = ops.sinkhorn(s, dummy_row=True, max_iter=20, tau=0.05)
    (y * cu(g[name + "_w"])).sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), g[name + "_out"], atol=2e-6)
    np.testing.assert_allclose(s.grad.cpu().numpy(), g[name + "_grad"], atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("shape", [(60, 60), (33, 57), (57, 33), (96, 96), (12, 90)])
def test_sinkhorn_small_bwd_vs_float64_autograd(shape):
    """The kernel's backward (recompute + reverse sweep) against float64 autograd of the oracle on the SAME fp32 input."""
    gen = torch.Generator().manual_seed(shape[0] * 100 + shape[1])
    s = torch.randn(*shape, generator=gen) * 2.0
    w = torch.randn(*shape, generator=gen)
    s64 = s.double().requires_grad_(True)
    (mgm_port.sinkhorn(s64, dummy_row=True, max_iter=20, tau=0.05) * w.double()).sum().backward()
    sc = s.cuda().requires_grad_(True)
    y = ops.sinkhorn(sc, dummy_row=True, max_iter=20, tau=0.05)
    (y * w.cuda()).sum().backward()
    ref = s64.grad.float().numpy()
    np.testing.assert_allclose(sc.grad.cpu().numpy(), ref, atol=2e-6 * np.abs(ref).max(), rtol=1e-5)


def test_sinkhorn_batched_projector_goldens(g):
    y = ops.sinkhorn(cu(g["skb_eq_le_in"]), dummy_row=True, max_iter=20, tau=0.1)
    np.testing.assert_allclose(y.cpu().numpy(), g["skb_eq_le_out"], atol=2e-6)
    v = cu(g["skb_eq_gt_in"])
    y = ops.sinkhorn(v.transpose(1, 2), dummy_row=True, max_iter=20, tau=0.1).transpose(1, 2)
    np.testing.assert_allclose(y.cpu().numpy(), g["skb_eq_gt_out"], atol=2e-6)
    for k in ("skb_rag", "skb_rag2"):
        y = ops.sinkhorn(cu(g[k + "_in"]), T(g[k + "_sizes"]), dummy_row=True, max_iter=20, tau=0.1)
        np.testing.assert_allclose(y.cpu().numpy(), g[k + "_out"], atol=2e-6)
    y = ops.sinkhorn(cu(g["sk50_in"]), max_iter=50, tau=0.05)
    np.testing.assert_allclose(y.cpu().numpy(), g["sk50_out"], atol=2e-6)
    y = ops.sinkhorn(cu(g["sk_tall_in"]), max_iter=20, tau=0.05)
    np.testing.assert_allclose(y.cpu().numpy(), g["sk_tall_out"], atol=2e-6)


def test_sinkhorn_edge_shapes():
    gen = torch.Generator().manual_seed(5)
    for shp, dummy in [((1, 1), True), ((1, 9), True), ((9, 1), True), ((96, 96), False), ((96, 95), True), ((2, 96), True)]:
        s = torch.randn(*shp, generator=gen)
        ref = mgm_port.sinkhorn(s.double(), dummy_row=dummy, max_iter=20, tau=0.05).float()
        out = ops.sinkhorn(s.cuda(), dummy_row=dummy, max_iter=20, tau=0.05).cpu()
        np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=1e-6)
    with pytest.raises(ValueError):
        ops.sinkhorn(torch.randn(4, device="cuda"))
    with pytest.raises(Exception):
        ops.sinkhorn(torch.randn(4, 4))            # CPU tensor: no fallback


@pytest.mark.parametrize("name", ["hung_40x32", "hung_20x32", "hung_32x32", "hung_23x57", "hung_ties"])
def test_hungarian_golden(g, name):
    out = ops.hungarian(cu(g[name + "_in"]))
    assert np.array_equal(out.cpu().numpy().astype(np.uint8), g[name + "_out"])


def test_hungarian_matches_scipy_restatement_random():
    rng = np.random.default_rng(1)
    mats = []
    for nr in range(1, 14):
        for nc in range(1, 14):
            mats.append(rng.standard_normal((nr, nc)).astype(np.float32))
            mats.append(rng.integers(0, 3, (nr, nc)).astype(np.float32))      # tie-heavy
            mats.append(np.full((nr, nc), 1.5, dtype=np.float32))             # constant
    for n in (23, 32, 40, 57, 90, 96, 128):
        mats.append(rng.standard_normal((n, 32)).astype(np.float32))
        mats.append(rng.standard_normal((32, n)).astype(np.float32))
        mats.append(rng.integers(0, 4, (n, 32)).astype(np.float32))
    # one padded batch launch
    N1, N2 = max(m.shape[0] for m in mats), max(m.shape[1] for m in mats)
    batch = np.zeros((len(mats), N1, N2), np.float32)
    for i, m in enumerate(mats):
        batch[i, :m.shape[0], :m.shape[1]] = m
    out = ops.hungarian(cu(batch), torch.tensor([m.shape[0] for m in mats]), torch.tensor([m.shape[1] for m in mats])).cpu().numpy()
    for i, m in enumerate(mats):
        ref = clib.hungarian(m)
        assert np.array_equal(out[i, :m.shape[0], :m.shape[1]], ref), (i, m.shape)
        assert out[i].sum() == min(m.shape)


def test_affinity_fwd_bwd_golden(g, sd):
    from adapteacher.modeling.GModule.utils.affinity import Affinity
    aff = Affinity(256).cuda()
    aff.load_state_dict({k[len("node_affinity."):]: v for k, v in sd.items() if k.startswith("node_affinity.")})
    X, Y = cu(g["aff_X"]).requires_grad_(True), cu(g["aff_Y"]).requires_grad_(True)
    M = aff(X, Y)
    np.testing.assert_allclose(M.detach().cpu().numpy(), g["aff_M"], atol=2e-6)
    (M * cu(g["aff_w"])).sum().backward()
    np.testing.assert_allclose(X.grad.cpu().numpy(), g["aff_dX"], atol=1e-6, rtol=1e-4)
    np.testing.assert_allclose(Y.grad.cpu().numpy(), g["aff_dY"], atol=1e-6, rtol=1e-4)
    for k, p in aff.named_parameters():
        gr = p.grad.cpu()
        # the golden gradients are fp32 sums with cancellation: tolerance relative to the tensor's scale
        if "aff_d_" + k in g.files:
            ref = g["aff_d_" + k]
            np.testing.assert_allclose(gr.numpy(), ref, atol=1e-5 * np.abs(ref).max(), rtol=1e-4, err_msg=k)
        else:
            sub = gr[::16, ::16] if gr.dim() == 2 else gr[::16]
            ref = g["aff_d_" + k + "_sub16"]
            np.testing.assert_allclose(sub.numpy(), ref, atol=1e-5 * np.abs(ref).max(), rtol=1e-4, err_msg=k)
            np.testing.assert_allclose(float(gr.double().sum()), float(g["aff_d_" + k + "_sum"]), rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(float((gr.double() ** 2).sum()), float(g["aff_d_" + k + "_sumsq"]), rtol=1e-4)


def test_affinity_degenerate_single_node(sd):
    from adapteacher.modeling.GModule.utils.affinity import Affinity
    aff = Affinity(256).cuda()
    out = aff(torch.randn(1, 256, device="cuda"), torch.randn(7, 256, device="cuda"))
    assert out.shape == (7,)                     # .squeeze() quirk (affinity.py:55)


def test_focal_bce_golden(g):
    from adapteacher.modeling.GModule.utils.losses import PermutationLoss
    S = cu(g["focal_S"]).requires_grad_(True)
    l = PermutationLoss()(S, cu(g["focal_Y"]), torch.tensor(23), torch.tensor(40))
    l.backward()
    np.testing.assert_allclose(l.item(), float(g["focal_loss"]), rtol=1e-6)
    np.testing.assert_allclose(S.grad.cpu().numpy(), g["focal_dS"], atol=1e-9, rtol=1e-5)
    with pytest.raises(AssertionError):
        PermutationLoss()(S.detach() + 2.0, cu(g["focal_Y"]))


def test_attention_adjacency_golden(g, sd):
    from adapteacher.modeling.GModule.utils.attentions import MultiHeadAttention
    att = MultiHeadAttention(256, 1, dropout=0.1, version="v2").cuda()
    att.load_state_dict({k[len("intra_domain_graph."):]: v for k, v in sd.items() if k.startswith("intra_domain_graph.")})
    x = cu(g["att_x"])
    n = x.shape[0]
    off = 1.0 - np.eye(n, dtype=np.float32)
    att.train()
    A = att.adjacency(x, [n], [cu(g["att_mask"])])
    np.testing.assert_allclose(A.cpu().numpy(), g["att_adj"] * off, atol=2e-7)
    att.eval()
    A = att.adjacency(x, [n])
    np.testing.assert_allclose(A.cpu().numpy(), g["att_adj_eval"] * off, atol=2e-7)
    _, full = att([x, x, x])
    np.testing.assert_allclose(full.cpu().numpy(), g["att_adj_eval"], atol=2e-7)
    # production dropout (Philox): ~10 % of the off-diagonal entries dropped, survivors scaled by 1/0.9, reproducible
    att.train()
    att.philox_offset = 0
    A1 = att.adjacency(x, [n]).cpu().numpy()
    att.philox_offset = 0
    A2 = att.adjacency(x, [n]).cpu().numpy()
    assert np.array_equal(A1, A2)
    ev = g["att_adj_eval"] * off
    kept = A1 != 0
    np.testing.assert_allclose(A1[kept], (ev / np.float32(0.9))[kept], rtol=1e-5)
    frac = 1.0 - kept.sum() / (n * n - n)
    assert 0.03 < frac < 0.2


def test_node_sampler_golden(golden_dir):
    gs = np.load(f"{golden_dir}/sampler.npz")
    from adapteacher.modeling.GModule.build_graph import PrototypeComputation

    class Boxes:
        def __init__(self, t):
            self.tensor = t

    class Inst:
        def __init__(self, b, c):
            self.pred_boxes, self.pred_classes = Boxes(b), c
            self._fields = {"pred_boxes": self.pred_boxes, "pred_classes": c}

        def __len__(self):
            return self.pred_boxes.tensor.shape[0]

    pc = PrototypeComputation(2, 10)
    for name, c in synth.SAMPLER_CASES.items():
        for channels_last in (False, True):
            feats = [f.cuda().requires_grad_(True) for f in synth.sampler_feats(name)]
            fin = [f.contiguous(memory_format=torch.channels_last) for f in feats] if channels_last else feats
            targets = [Inst(torch.tensor(b, dtype=torch.float32).reshape(-1, 4).cuda(), torch.tensor(k, dtype=torch.int64).cuda())
                       for b, k in zip(c["boxes"], c["classes"])]
            nodes, labels = pc(fin, targets)
            assert len(nodes) == int(gs[name + "_nout"])
            for i, (n, l) in enumerate(zip(nodes, labels)):
                assert np.array_equal(n.detach().cpu().numpy(), gs[f"{name}_nodes{i}"]), (name, i)
                assert np.array_equal(l.cpu().numpy(), gs[f"{name}_labels{i}"]), (name, i)
            # backward = scatter of the node gradients into the pyramid (autograd of the reference's indexing)
            w = [torch.randn_like(n) for n in nodes]
            sum((n * wi).sum() for n, wi in zip(nodes, w)).backward()
            feats_cpu = [f.detach().cpu().requires_grad_(True) for f in feats]
            rn, _ = mgm_port.sample_nodes(feats_cpu, [t.pred_boxes.tensor.cpu() for t in targets],
                                          [t.pred_classes.cpu() for t in targets])
            sum((n * wi.cpu()).sum() for n, wi in zip(rn, w)).backward()
            for f, fc in zip(feats, feats_cpu):
                assert torch.equal(f.grad.cpu(), fc.grad)
    assert pc([f.cuda() for f in synth.sampler_feats("samp_a")],
              [Inst(torch.zeros(0, 4).cuda(), torch.zeros(0, dtype=torch.int64).cuda())] * 2) == (None, None)


@pytest.mark.parametrize("N,B", [(128, 3), (256, 2), (512, 2), (1024, 1), (100, 2)])
def test_sinkhorn_large_vs_oracle(N, B):
    gen = torch.Generator().manual_seed(N)
    s = torch.randn(B, N, N, generator=gen)
    ref = mgm_port.sinkhorn(s.double(), max_iter=50, tau=0.05).float()
    out = ops.sinkhorn(s.cuda(), max_iter=50, tau=0.05).cpu()
    # fp32 path: x / tau has an ulp of 8e-6 at |x / tau| ~ 64-128, so agreement is ~1e-5 relative
    np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=2e-5, rtol=2e-4)
    np.testing.assert_allclose(out.sum(1).numpy(), 1.0, atol=1e-4)          # last step normalises columns


def test_sinkhorn_large_full_size_properties():
    """BASELINE.json configs[4] sizes: properties that do not need the CPU oracle."""
    for N, B in [(1024, 16), (512, 40), (256, 150)]:
        s = torch.randn(B, N, N, device="cuda")
        out = ops.sinkhorn(s, max_iter=50, tau=0.05)
        assert torch.isfinite(out).all() and (out >= 0).all() and (out <= 1 + 1e-5).all()
        torch.testing.assert_close(out.sum(1), torch.ones(B, N, device="cuda"), atol=2e-4, rtol=0)
        # invariance: a constant added to a row of the input is removed by the first (row) normalisation
        s2 = s + torch.randn(B, N, 1, device="cuda")
        torch.testing.assert_close(ops.sinkhorn(s2, max_iter=50, tau=0.05), out, atol=3e-4, rtol=1e-2)
        # batch items are independent
        torch.testing.assert_close(ops.sinkhorn(s[B // 2:B // 2 + 1], max_iter=50, tau=0.05), out[B // 2:B // 2 + 1], atol=0, rtol=0)


@pytest.mark.parametrize("n1,n2", [(256, 256), (300, 512)])
def test_affinity_large_vs_oracle(n1, n2):
    """BASELINE.json configs[4] sizes (the microbench runs N = 256 / 512 / 1024): the separable-form affinity beyond the
    96-node graphs of the matching stage, forward and every gradient, against the reference-form oracle (which materialises
    the n1 x n2 x 512 tensor, utils/affinity.py:48-54)."""
    from adapteacher.modeling.GModule.utils.affinity import Affinity
    sd = synth.perturb_affinity_state(synth.mgm_unsup_state(0), 0)
    aff = Affinity(256).cuda()
    aff.load_state_dict({k[len("node_affinity."):]: v for k, v in sd.items() if k.startswith("node_affinity.")})
    g = torch.Generator().manual_seed(31)
    X, Y, up = torch.randn(n1, 256, generator=g), torch.randn(n2, 256, generator=g), torch.randn(n1, n2, generator=g)
    sdc = {k: v.clone().double().requires_grad_(k.startswith("node_affinity.")) for k, v in sd.items()}
    Xc, Yc = X.clone().double().requires_grad_(True), Y.clone().double().requires_grad_(True)
    ref = mgm_port.affinity(sdc, Xc, Yc)
    (ref * up.double()).sum().backward()
    Xg, Yg = X.cuda().requires_grad_(True), Y.cuda().requires_grad_(True)
    out = aff(Xg, Yg)
    (out * up.cuda()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), atol=2e-6 * float(ref.abs().max()), rtol=1e-5)
    rel = lambda a, r: float((a.double().cpu() - r).norm() / r.norm())
    assert rel(Xg.grad, Xc.grad) < 1e-5 and rel(Yg.grad, Yc.grad) < 1e-5
    for name, p in (("fc_M.0.weight", aff.fc_M[0].weight), ("fc_M.2.weight", aff.fc_M[2].weight),
                    ("project_sr.weight", aff.project_sr.weight), ("project_tg.weight", aff.project_tg.weight)):
        assert rel(p.grad, sdc["node_affinity." + name].grad) < 1e-5, name
