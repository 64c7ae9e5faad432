"""PseudoLabRPN; mirrors reference adapteacher/modeling/proposal_generator/rpn.py:10-55 (the d2 RPN with a
``compute_loss`` switch so proposals can be produced in train mode without ground truth).  Inference form only."""
from ttdg_b200.detector import RPN as PseudoLabRPN  # noqa: F401
