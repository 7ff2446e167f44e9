"""Oracle restatements of the graph-matching operators (rows a4-a8, a13 of SURVEY.md §8).
Plain fp32 PyTorch on whatever device the inputs live on (CPU in the tests).  Test infrastructure.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- a4: Affinity
def affinity(X, Y, p, prefix=""):
    """models/affinity_layer.py:52-73.  p: project_sr.weight, project_tg.weight,
    fc_M.0.{weight,bias}, fc_M.2.{weight,bias}.  Literal (non-separable) form: the pairwise
    [N1,N2,512] tensor is materialised exactly as the reference does."""
    xs = X @ p[prefix + "project_sr.weight"].t()
    yt = Y @ p[prefix + "project_tg.weight"].t()
    n1, n2 = xs.shape[0], yt.shape[0]
    pair = torch.cat([xs[:, None, :].expand(n1, n2, -1), yt[None, :, :].expand(n1, n2, -1)], dim=-1)
    hid = torch.relu(F.linear(pair, p[prefix + "fc_M.0.weight"], p[prefix + "fc_M.0.bias"]))
    return F.linear(hid, p[prefix + "fc_M.2.weight"], p[prefix + "fc_M.2.bias"]).squeeze()


# ----------------------------------------------------------------------------- a5: Sinkhorn (RPM)
def instance_norm_matrix(M, eps=1e-5):
    """nn.InstanceNorm2d(1) applied to M[None, None] (graph_matching.py:177, 574): whole-matrix
    mean / biased variance, no affine."""
    mu = M.mean()
    var = M.var(unbiased=False)
    return (M - mu) / torch.sqrt(var + eps)


def sinkhorn_rpm(log_alpha, n_iters=5, slack=True):
    """graph_matching.py:637-689 (eps<0 branch).  log_alpha [B,J,K]."""
    if slack:
        b, j, k = log_alpha.shape
        pad = log_alpha.new_zeros(b, j + 1, k + 1)
        pad[:, :j, :k] = log_alpha
        cur = pad
        for _ in range(n_iters):
            top = cur[:, :-1, :] - torch.logsumexp(cur[:, :-1, :], dim=2, keepdim=True)
            cur = torch.cat([top, cur[:, -1:, :]], dim=1)                       # :661-664
            left = cur[:, :, :-1] - torch.logsumexp(cur[:, :, :-1], dim=1, keepdim=True)
            cur = torch.cat([left, cur[:, :, -1:]], dim=2)                       # :666-669
        return cur[:, :-1, :-1]
    cur = log_alpha
    for _ in range(n_iters):
        cur = cur - torch.logsumexp(cur, dim=2, keepdim=True)
        cur = cur - torch.logsumexp(cur, dim=1, keepdim=True)
    return cur


def sinkhorn_rpm_exp(M, n_iters=20, instnorm=True):
    """InstNorm_layer -> sinkhorn_rpm -> exp, i.e. graph_matching.py:574-575.  M [N1,N2]."""
    z = instance_norm_matrix(M) if instnorm else M
    return sinkhorn_rpm(z[None], n_iters=n_iters, slack=True)[0].exp()


def bce_focal(prob, target, gamma=2.0, alpha=0.25):
    """BCEFocalLoss.forward with reduction='elementwise_mean' (graph_matching.py:31-38)."""
    loss = -alpha * (1 - prob) ** gamma * target * torch.log(prob) \
           - (1 - alpha) * prob ** gamma * (1 - target) * torch.log(1 - prob)
    return loss.mean()


def matching_loss_o2o(Mn, labels_1, labels_2, num_classes):
    """TP/FP focal losses on the Sinkhorn-normalised matrix (graph_matching.py:572-590)."""
    eye = torch.eye(num_classes, device=Mn.device)
    target = eye[labels_1.long()] @ eye[labels_2.long()].t()
    tp_mask = (target == 1).float()
    idx = (Mn * tp_mask).max(-1)[1]
    tp = Mn[torch.arange(Mn.shape[0], device=Mn.device), idx].view(-1, 1)
    fp = Mn[target == 0].view(-1, 1)
    tp_loss = bce_focal(tp, torch.ones_like(tp)) / len(tp)
    fp_loss = bce_focal(fp, torch.zeros_like(fp)) / fp.sum().detach()
    return tp_loss + fp_loss


def forward_aff(nodes_1, nodes_2, labels_1, labels_2, p, num_classes, prefix="node_affinity."):
    """GModule._forward_aff, 'o2o' branch (graph_matching.py:569-590, 599)."""
    M = affinity(nodes_1, nodes_2, p, prefix)
    Mn = sinkhorn_rpm_exp(M, 20, True)
    return matching_loss_o2o(Mn, labels_1, labels_2, num_classes), Mn


def forward_qu(edge_1, edge_2, Mn):
    """GModule._forward_qu (graph_matching.py:604-607): mean |E1 M - M E2|."""
    return (edge_1 @ Mn - Mn @ edge_2).abs().mean()


# ----------------------------------------------------------------------------- a7: attention
def mha_v2(key, value, query, p, prefix="", dropout=0.0, training=False):
    """MultiHeadAttention.forward, version='v2', num_heads=1 (models/transformer.py:43-75, 110)
    with dot_attention (:13-23).  Inputs [N,256]; returns (LayerNorm(query + proj), attention).
    dropout is applied only when training (parity tests run p=0 / eval)."""
    d = p[prefix + "linear_k.weight"].shape[0]
    k = F.linear(key, p[prefix + "linear_k.weight"], p[prefix + "linear_k.bias"])
    v = F.linear(value, p[prefix + "linear_v.weight"], p[prefix + "linear_v.bias"])
    q = F.linear(query, p[prefix + "linear_q.weight"], p[prefix + "linear_q.bias"])
    scale = float(d) ** -0.5                       # (key.size(-1) // num_heads) ** -0.5, heads = 1
    att = torch.softmax((q @ k.t()) * scale, dim=-1)
    att = F.dropout(att, dropout, training)
    ctx = att @ v
    out = F.linear(ctx, p[prefix + "linear_final.weight"], p[prefix + "linear_final.bias"])
    out = F.dropout(out, dropout, training)
    out = F.layer_norm(query + out, (d,), p[prefix + "layer_norm.weight"], p[prefix + "layer_norm.bias"])
    return out.squeeze(), att.squeeze()


# ----------------------------------------------------------------------------- a8: node heads
def ln_mlp(x, p, prefix, layers, final_ln):
    """Linear -> LayerNorm(no affine) -> ReLU chains: head_in_ln (graph_matching.py:148-154,
    Linear indices 0,3 with a trailing LN) and node_dis_2 (:191-202, indices 0,3,6,9)."""
    for n, li in enumerate(layers):
        x = F.linear(x, p[f"{prefix}{li}.weight"], p[f"{prefix}{li}.bias"])
        last = n == len(layers) - 1
        if not last or final_ln:
            x = F.layer_norm(x, (x.shape[-1],))
        if not last:
            x = torch.relu(x)
    return x


def head_in_ln(x, p, prefix="head_in_ln."):
    return ln_mlp(x, p, prefix, (0, 3), final_ln=True)


def node_dis(x, p, prefix="node_dis_2."):
    return ln_mlp(x, p, prefix, (0, 3, 6, 9), final_ln=False)


def node_cls(x, p, prefix="node_cls_middle."):
    """node_cls_middle (graph_matching.py:158-162)."""
    h = torch.relu(F.linear(x, p[prefix + "0.weight"], p[prefix + "0.bias"]))
    return F.linear(h, p[prefix + "2.weight"], p[prefix + "2.bias"])


# ----------------------------------------------------------------------------- a13: SinkhornDistance
def sinkhorn_distance(x, y, eps, max_iter, reduction="none", thresh=0.1):
    """utils/sinkhorn_distance.py:27-86.  Returns (cost, pi, C, executed_iterations)."""
    C = ((x.unsqueeze(-2) - y.unsqueeze(-3)).abs() ** 2).sum(-1)                 # :81-86
    p1, p2 = x.shape[-2], y.shape[-2]
    bsz = 1 if x.dim() == 2 else x.shape[0]
    mu = torch.full((bsz, p1), 1.0 / p1, dtype=torch.float, device=C.device).squeeze()
    nu = torch.full((bsz, p2), 1.0 / p2, dtype=torch.float, device=C.device).squeeze()
    u, v = torch.zeros_like(mu), torch.zeros_like(nu)

    def cost_m(u_, v_):                                                           # :75-78
        return (-C + u_.unsqueeze(-1) + v_.unsqueeze(-2)) / eps

    nits = 0
    for _ in range(max_iter):
        u_prev = u
        u = eps * (torch.log(mu + 1e-8) - torch.logsumexp(cost_m(u, v), dim=-1)) + u
        v = eps * (torch.log(nu + 1e-8) - torch.logsumexp(cost_m(u, v).transpose(-2, -1), dim=-1)) + v
        err = (u - u_prev).abs().sum(-1).mean()
        nits += 1
        if err.item() < thresh:
            break
    pi = torch.exp(cost_m(u, v))
    cost = (pi * C).sum((-2, -1))
    if reduction == "mean":
        cost = cost.mean()
    elif reduction == "sum":
        cost = cost.sum()
    return cost, pi, C, nits
