"""Mirror of utils/sinkhorn_distance.py: SinkhornDistance(eps, max_iter, reduction).forward(x, y)
-> (cost, pi, C), on the sm_100a kernels ge_sinkhorn_distance_fwd/bwd (one CTA per batch element,
no per-iteration host sync; the batch-mean early stop is reproduced on the device)."""
import torch
from torch import nn

from .. import functional as GF


class SinkhornDistance(nn.Module):
    def __init__(self, eps, max_iter, reduction="none"):
        super().__init__()
        self.eps = eps
        self.max_iter = max_iter
        self.reduction = reduction
        self.thresh = 1e-1          # sinkhorn_distance.py:49

    def forward(self, x, y):
        two_d = x.dim() == 2
        xb, yb = (x[None], y[None]) if two_d else (x, y)
        with torch.autocast("cuda", enabled=False):
            cost, pi, C, nits = GF.sinkhorn_distance(xb.float(), yb.float(), self.eps, self.max_iter, self.thresh)
        self.last_iterations = nits          # device int32 [1]; read it only if you accept a sync
        if two_d:
            cost, pi, C = cost[0], pi[0], C[0]
        if self.reduction == "mean":
            cost = cost.mean()
        elif self.reduction == "sum":
            cost = cost.sum()
        return cost, pi, C

    def M(self, C, u, v):
        """(-C + u_i + v_j) / eps  (sinkhorn_distance.py:75-78)."""
        return (-C + u.unsqueeze(-1) + v.unsqueeze(-2)) / self.eps

    @staticmethod
    def _cost_matrix(x, y, p=2):
        return torch.sum(torch.abs(x.unsqueeze(-2) - y.unsqueeze(-3)) ** p, -1)

    @staticmethod
    def ave(u, u1, tau):
        return tau * u + (1 - tau) * u1
