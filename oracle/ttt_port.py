"""CPU restatement of one test-time-adaptation step + the eval pass (adapteacher/engine/trainer.py:469-485,
meta_arch/rcnn.py:331-357) assembled from detector_port (d2 restatement, parity unpinned) and mgm_port (pinned to
the reference's own MGM modules).  TEST INFRASTRUCTURE: the checker in tests / smoke and the timed `cpu_baseline` /
`--impl reference` leg of bench.py.

``dtype=torch.float64`` runs the same two passes in float64 (weights and images widened exactly): the noise-free limit of
the algorithm.  The detector has discontinuities (top-k, NMS, score threshold, ReLU masks in the backward), so the parity
tests measure the CUDA path AND the fp32 restatement against that limit instead of against each other only."""
import contextlib

import torch

from oracle import detector_port as dp
from oracle import mgm_port

ADAPTED_PREFIXES = ("backbone.bottom_up.res3", "backbone.bottom_up.res4", "backbone.bottom_up.res5", "backbone.fpn_")


def adapted_keys(sd_det, sd_mgm):
    det = [k for k in sd_det if k.startswith(ADAPTED_PREFIXES) and ".norm." not in k]
    mgm = [k for k in sd_mgm if k.startswith("node_affinity.")]
    return det, mgm


class Trainer:
    """Holds leaf parameters and a torch.optim.SGD like the reference's caller (lr .005, momentum .9, wd 1e-4)."""

    def __init__(self, sd_det, sd_mgm, U, lr=0.005, momentum=0.9, weight_decay=1e-4, dtype=torch.float32):
        self.dtype = dtype
        self.sd_det = {k: v.clone().to(dtype) for k, v in sd_det.items()}
        self.sd_mgm = {k: v.clone().to(dtype) for k, v in sd_mgm.items()}
        self.U = U.clone().to(dtype)
        self.det_keys, self.mgm_keys = adapted_keys(sd_det, sd_mgm)
        for k in self.det_keys:
            self.sd_det[k].requires_grad_(True)
        for k in self.mgm_keys:
            self.sd_mgm[k].requires_grad_(True)
        self.params = [self.sd_det[k] for k in self.det_keys] + [self.sd_mgm[k] for k in self.mgm_keys]
        self.opt = torch.optim.SGD(self.params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.last = None

    def _ctx(self):
        return dp.float64() if self.dtype == torch.float64 else contextlib.nullcontext()

    def named_adapted(self):
        """(d2 / reference state-dict key, parameter) of everything the test-time loss reaches."""
        return [(k, self.sd_det[k]) for k in self.det_keys] + \
               [("multi_matching_unsup." + k, self.sd_mgm[k]) for k in self.mgm_keys]

    def ttt_step(self, images_u8, keep_masks=None, precise=False, dets_override=None, U_override=None, step=True):
        """One adaptation step.  Parity hooks (tests only): ``dets_override`` = per image (boxes, scores, classes) fed to
        the node sampler instead of this path's own box-head detections (the discrete part of the detector),
        ``U_override`` = the matching result (the chaotic part), ``keep_masks`` = dropout keep-masks
        (list, or a callable of the graph sizes).  ``self.last`` keeps the intermediate results."""
        with self._ctx():
            feats, props, dets = dp.forward_ttt(self.sd_det, images_u8)
            use = dets if dets_override is None else dets_override
            nodes, labels = mgm_port.sample_nodes(feats, [d[0].float() for d in use], [d[2] for d in use])
            self.last = {"feats": feats, "props": props, "dets": dets, "nodes": nodes}
            if nodes is None:
                return None
            if callable(keep_masks):
                keep_masks = keep_masks([int(n.shape[0]) for n in nodes])
            out = mgm_port.mgm3_unsup_forward(self.sd_mgm, nodes, labels, self.U, keep_masks, precise=precise,
                                              U_override=U_override, return_aux=True)
            if out is None:
                return None
            loss, aux = out
            self.last.update(aux=aux, loss=loss.detach())
            self.opt.zero_grad()
            loss.backward()
            if step:
                self.opt.step()
        return float(loss.detach())

    def eval_pass(self, images_u8, full=False):
        with self._ctx():
            sd = {k: v.detach() for k, v in self.sd_det.items()}
            out = dp.inference(sd, images_u8)
        return out if full else out[0]
