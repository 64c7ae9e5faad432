"""pad_tensor; mirrors reference adapteacher/modeling/GModule/utils/pad_tensor.py:5-31 (zero-pad a list of
tensors to their common maximum shape).  The GA-GM solver here is ragged-native and does not need it."""
import torch
import torch.nn.functional as F


def pad_tensor(inp):
    assert type(inp[0]) == torch.Tensor
    max_shape = [max(t.shape[d] for t in inp) for d in range(inp[0].dim())]
    out = []
    for t in inp:
        pad = []
        for d in reversed(range(t.dim())):
            pad += [0, max_shape[d] - t.shape[d]]
        out.append(F.pad(t, pad, 'constant', 0))
    return out
