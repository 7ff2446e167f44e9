"""Mirror of models/gradient_reversal.py: GradientReversalFunction / GradientReversal / FocalLoss.

The reference's forward does a full `x.clone()` (gradient_reversal.py:18); here the forward is a
view and only the backward does work (-lambda * grad), which XLA-free eager autograd fuses into
the consumer's grad stream."""
import torch
import torch.nn.functional as F

from ..functional import _GradReverse


class GradientReversalFunction(_GradReverse):
    """apply(x, lambda_) -> x ; backward: -lambda_ * grad  (gradient_reversal.py:6-24)."""


class GradientReversal(torch.nn.Module):
    def __init__(self, lambda_=1):
        super().__init__()
        self.lambda_ = lambda_

    def forward(self, x):
        return GradientReversalFunction.apply(x, self.lambda_)

    def extra_repr(self):
        return f"lambda_={self.lambda_}"


def FocalLoss(inputs, targets, gamma=5.0):
    """gradient_reversal.py:33-37 (unused by the trainers; API surface)."""
    bce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    return ((1 - torch.exp(-bce)) ** gamma * bce).mean()
