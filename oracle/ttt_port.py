"""CPU restatement of one test-time-adaptation step + the eval pass (adapteacher/engine/trainer.py:469-485,
meta_arch/rcnn.py:331-357) assembled from detector_port (d2 restatement, parity unpinned) and mgm_port (pinned to
the reference's own MGM modules).  TEST INFRASTRUCTURE: the checker in tests / smoke and the timed `cpu_baseline` /
`--impl reference` leg of bench.py."""
import torch

from oracle import detector_port as dp
from oracle import mgm_port

ADAPTED_PREFIXES = ("backbone.bottom_up.res3", "backbone.bottom_up.res4", "backbone.bottom_up.res5", "backbone.fpn_")


def adapted_keys(sd_det, sd_mgm):
    det = [k for k in sd_det if k.startswith(ADAPTED_PREFIXES) and ".norm." not in k]
    mgm = [k for k in sd_mgm if k.startswith("node_affinity.")]
    return det, mgm


class Trainer:
    """Holds fp32 leaf parameters and a torch.optim.SGD like the reference's caller (lr .005, momentum .9, wd 1e-4)."""

    def __init__(self, sd_det, sd_mgm, U, lr=0.005, momentum=0.9, weight_decay=1e-4):
        self.sd_det = {k: v.clone() for k, v in sd_det.items()}
        self.sd_mgm = {k: v.clone() for k, v in sd_mgm.items()}
        self.U = U.clone()
        det, mgm = adapted_keys(sd_det, sd_mgm)
        for k in det:
            self.sd_det[k].requires_grad_(True)
        for k in mgm:
            self.sd_mgm[k].requires_grad_(True)
        self.params = [self.sd_det[k] for k in det] + [self.sd_mgm[k] for k in mgm]
        self.opt = torch.optim.SGD(self.params, lr=lr, momentum=momentum, weight_decay=weight_decay)

    def ttt_step(self, images_u8, keep_masks=None, precise=False):
        feats, props, dets = dp.forward_ttt(self.sd_det, images_u8)
        nodes, labels = mgm_port.sample_nodes(feats, [d[0] for d in dets], [d[2] for d in dets])
        if nodes is None:
            return None
        loss = mgm_port.mgm3_unsup_forward(self.sd_mgm, nodes, labels, self.U, keep_masks, precise=precise)
        if loss is None:
            return None
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def eval_pass(self, images_u8):
        sd = {k: v.detach() for k, v in self.sd_det.items()}
        return dp.inference(sd, images_u8)[0]
