"""hungarian(); mirrors reference adapteacher/modeling/GModule/utils/hungarian.py:8-65 (SciPy
``linear_sum_assignment`` on the host) with an on-device fp64 LAP that keeps SciPy's tie-breaking."""
from ttdg_b200 import ops


def hungarian(s, n1=None, n2=None, nproc=1):
    """``nproc`` is accepted for signature compatibility; the batch is solved one warp per matrix."""
    return ops.hungarian(s, n1, n2)
