"""Mirror of models/affinity_layer.py: Affinity(d).forward(X, Y) -> M [N1, N2].

state_dict keys are the reference's (fc_M.0/2, project_sr, project_tg).  The forward never
builds the [N1,N2,512] pair tensor (affinity_layer.py:60-63): fc_M.0 is linear, so
fc_M.0([xs_i ; yt_j]) = A_i + B_j with two small dense projections (cuBLAS), and the
relu-coupled pairwise reduction runs in the sm_100a kernel ge_affinity_pairwise_fwd/bwd."""
import torch
from torch import nn

from .. import functional as GF


class Affinity(nn.Module):
    def __init__(self, d=256):
        super().__init__()
        self.d = d
        self.fc_M = nn.Sequential(nn.Linear(2 * d, 2 * d), nn.ReLU(), nn.Linear(2 * d, 1))
        self.project_sr = nn.Linear(d, d, bias=False)
        self.project_tg = nn.Linear(d, d, bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        # affinity_layer.py:35-43: N(0, 0.01) weights, zero biases
        for layer in (self.fc_M[0], self.fc_M[2]):
            nn.init.normal_(layer.weight, std=0.01)
            nn.init.constant_(layer.bias, 0)
        nn.init.normal_(self.project_sr.weight, std=0.01)
        nn.init.normal_(self.project_tg.weight, std=0.01)

    def pair_operands(self, X, Y):
        """A [N1,2d], B [N2,2d] with fc_M.0([xs;yt]) == A_i + B_j."""
        d = self.d
        W1, b1 = self.fc_M[0].weight, self.fc_M[0].bias
        A = self.project_sr(X) @ W1[:, :d].t()
        B = torch.addmm(b1, self.project_tg(Y), W1[:, d:].t())
        return A, B

    def forward(self, X, Y):
        with torch.autocast("cuda", enabled=False):
            A, B = self.pair_operands(X.float(), Y.float())
            M = GF.affinity_pairwise(A, B, self.fc_M[2].weight.view(-1), self.fc_M[2].bias)
        return M.squeeze()
