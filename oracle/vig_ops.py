"""Oracle restatements of the ViG graph operators (rows a9-a11 of SURVEY.md §8): k-NN graph
build, max-relative graph conv, Grapher.  Plain fp32 PyTorch.  Test infrastructure."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def pairwise_sq_dist(x, y):
    """(xy_)pairwise_distance (models/vig.py:232-274): x [B,N,C], y [B,M,C] ->
    (|x|^2 + (-2 x.y^T)) + |y|^2^T, in that association order."""
    inner = -2 * torch.matmul(x, y.transpose(2, 1))
    xs = (x * x).sum(-1, keepdim=True)
    ys = (y * y).sum(-1, keepdim=True)
    return xs + inner + ys.transpose(2, 1)


def knn_distances(x, y=None, relative_pos=None):
    """The distance matrix torch.topk sees in DenseDilatedKnnGraph.forward (vig.py:369-381):
    channel-wise L2 normalisation first (:372-373/378), then the pairwise form.
    x [B,C,N,1], y [B,C,M,1] or None.  Returns [B,N,M]."""
    xn = F.normalize(x, p=2.0, dim=1)
    yn = F.normalize(y, p=2.0, dim=1) if y is not None else xn
    xt = xn.transpose(2, 1).squeeze(-1)
    yt = yn.transpose(2, 1).squeeze(-1)
    dist = pairwise_sq_dist(xt, yt)
    if relative_pos is not None:
        dist = dist + relative_pos
    return dist


def dense_dilated_knn(x, y=None, k=9, dilation=1, relative_pos=None):
    """DenseDilatedKnnGraph.forward -> edge_index int64 [2,B,N,k] (vig.py:277-329, 344-354, 369-381).
    [0] neighbour index (ascending distance), [1] centre index."""
    with torch.no_grad():
        dist = knn_distances(x, y, relative_pos)
        b, n, _ = dist.shape
        _, nn_idx = torch.topk(-dist, k=k * dilation)
        centre = torch.arange(n, device=dist.device).view(1, n, 1).expand(b, n, k * dilation)
        edge = torch.stack((nn_idx, centre), dim=0)
    return edge[:, :, :, ::dilation]


def gather_points(x, idx):
    """batched_index_select (vig.py:209-229): x [B,C,M,1], idx [B,N,k] -> [B,C,N,k]."""
    b, c, m = x.shape[:3]
    _, n, k = idx.shape
    flat = x.squeeze(-1).transpose(1, 2).reshape(b * m, c)
    sel = flat[(idx + torch.arange(b, device=idx.device).view(-1, 1, 1) * m).reshape(-1)]
    return sel.view(b, n, k, c).permute(0, 3, 1, 2)


def max_relative(x, edge_index, y=None):
    """First half of MRConv2d.forward (vig.py:96-104): channel-interleaved [x ; max_k(x_j - x_i)]
    as [B,2C,N,1]."""
    x_i = gather_points(x, edge_index[1])
    x_j = gather_points(y if y is not None else x, edge_index[0])
    rel = (x_j - x_i).max(-1, keepdim=True)[0]
    b, c, n, _ = x.shape
    return torch.stack([x, rel], dim=2).reshape(b, 2 * c, n, 1)


def basic_conv(x, p, prefix, norm=None, act="gelu", training=True, groups=4):
    """BasicConv([cin, cout]) (vig.py:476-488): grouped 1x1 conv (+BN) (+act)."""
    x = F.conv2d(x, p[prefix + "0.weight"], p.get(prefix + "0.bias"), groups=groups)
    if norm == "batch":
        x = F.batch_norm(x, p[prefix + "1.running_mean"], p[prefix + "1.running_var"],
                         p[prefix + "1.weight"], p[prefix + "1.bias"], training, 0.1, 1e-5)
    if act == "gelu":
        x = F.gelu(x)
    elif act == "relu":
        x = torch.relu(x)
    return x


def mrconv(x, edge_index, p, prefix, y=None, norm=None, act="gelu", training=True):
    """MRConv2d.forward (vig.py:96-105); prefix addresses `...gconv.nn.`."""
    return basic_conv(max_relative(x, edge_index, y), p, prefix, norm, act, training)


def conv_bn(x, p, prefix, training=True):
    """nn.Sequential(Conv2d 1x1, BatchNorm2d) as in Grapher.fc1 / fc2 (vig.py:394-403)."""
    x = F.conv2d(x, p[prefix + "0.weight"], p[prefix + "0.bias"])
    return F.batch_norm(x, p[prefix + "1.running_mean"], p[prefix + "1.running_var"],
                        p[prefix + "1.weight"], p[prefix + "1.bias"], training, 0.1, 1e-5)


def grapher(x, p, prefix="", k=9, dilation=1, r=1, norm="batch", act="gelu", training=True,
            relative_pos=None, return_edges=False):
    """Grapher.forward with conv='mr' (vig.py:422-430) over vig.DyGraphConv2d.forward (:196-206).
    x [B,C,H,W]."""
    b, c, h, w = x.shape
    t = conv_bn(x, p, prefix + "fc1.", training)
    y = None
    if r > 1:
        y = F.avg_pool2d(t, r, r).reshape(b, c, -1, 1)
    tt = t.reshape(b, c, -1, 1)
    edge = dense_dilated_knn(tt, y, k, dilation, relative_pos)
    g = mrconv(tt, edge, p, prefix + "graph_conv.gconv.nn.", y, norm, act, training)
    g = g.reshape(b, -1, h, w)
    out = conv_bn(g, p, prefix + "fc2.", training) + x
    return (out, edge) if return_edges else out
