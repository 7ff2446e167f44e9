"""Oracle restatement of TGCN.forward and its DyGraphConv2d (row a12 of SURVEY.md §8),
functional over a reference-keyed state dict.  Plain fp32 PyTorch.  Test infrastructure."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import graph_ops as G
from . import vig_ops as V
from .fpn_ops import grad_reverse


def tgcn_dygraph_step(levels, rs, hidden, pos, p, training=True, dropout=0.1, return_edges=False):
    """TGCN.DyGraphConv2d.forward (models/TGCN.py:62-78): pool+concat -> MLP -> +pos -> kNN(x, hidden)
    -> MRConv(x, idx, hidden).  levels: list of [B,C,s,s]; hidden [B,C,N]."""
    pooled = [F.avg_pool2d(t, r, r) if r > 1 else t for t, r in zip(levels, rs)]
    x = torch.cat(pooled, dim=1)
    x = F.conv2d(x, p["grapher.MLP.0.weight"], p["grapher.MLP.0.bias"])
    x = F.batch_norm(x, p["grapher.MLP.1.running_mean"], p["grapher.MLP.1.running_var"],
                     p["grapher.MLP.1.weight"], p["grapher.MLP.1.bias"], training, 0.1, 1e-5)
    x = F.dropout(F.gelu(x), dropout, training)
    x = F.conv2d(x, p["grapher.MLP.4.weight"], p["grapher.MLP.4.bias"])
    x = x + pos
    b, c, h, w = x.shape
    x = x.reshape(b, c, -1, 1)
    edge = V.dense_dilated_knn(x, hidden, 9, 1)                                   # TGCN.py:76 (k=9, d=1)
    out = V.mrconv(x, edge, p, "grapher.gconv.nn.", y=hidden, norm=None, act="gelu", training=training)
    out = out.reshape(b, -1, h * w)
    return (out, edge) if return_edges else out


def tgcn_forward(features, nodes, p, rs=(8, 4, 2, 1), clip_hw=(8, 8), transport="node_discriminate",
                 sinkhorn=None, training=True, dropout=0.1, return_debug=False):
    """TGCN.forward (models/TGCN.py:224-285), cluster_method=None.
    features: 4 x [b,t,C,s,s]; nodes: (source_nodes [Ns,256], target_nodes [Nt,256])."""
    f1 = features[0]
    b, t = f1.shape[:2]
    c = p["pos_embed"].shape[2]
    hidden = torch.zeros(b, c, clip_hw[0] * clip_hw[1]).type_as(f1)               # :230
    edges = []
    for i in range(t):
        lv = [f[:, i] for f in features]
        hidden, e = tgcn_dygraph_step(lv, rs, hidden, p["pos_embed"][i], p, training, dropout, True)
        edges.append(e)
    cur = hidden
    # prediction head (:184-190, 238-239): result unused without a cluster method, but BN stats move
    of = cur.reshape(b, c, clip_hw[0], clip_hw[1])
    of = F.conv2d(of, p["prediction.0.weight"], p["prediction.0.bias"], stride=2)
    of = F.batch_norm(of, p["prediction.1.running_mean"], p["prediction.1.running_var"],
                      p["prediction.1.weight"], p["prediction.1.bias"], training, 0.1, 1e-5)
    of = F.adaptive_avg_pool2d(F.dropout(F.gelu(of), dropout, training), 1).view(b, -1)
    src, tgt = nodes
    og = cur.transpose(1, 2)                                                      # [b, N, C]
    bg, dg, ng = og.shape
    flat = og.reshape(bg * dg, ng)
    allnodes = torch.cat([flat, src, tgt])
    att = G.mha_v2(allnodes, allnodes, allnodes, p, "graph_attention.", dropout, training)[0]     # :265-266
    nodes_g = att[: bg * dg].reshape(bg, dg, ng)
    losses = {}
    if transport == "node_discriminate":                                          # :272-279
        ns = nodes_g[: bg // 2].reshape(-1, ng)
        nt = nodes_g[bg // 2:].reshape(-1, ng)
        rev = grad_reverse(torch.cat([ns, nt], dim=0), 0.02)
        tg = torch.cat([torch.ones(len(ns)), torch.zeros(len(nt))]).to(rev.device)
        losses["node_dis_loss"] = 0.1 * F.binary_cross_entropy_with_logits(G.node_dis(rev, p).view(-1), tg)
    elif transport == "sinkhorn_distance":                                        # :281-283
        eps, max_iter, red = sinkhorn
        losses["sinkhorn_loss"] = G.sinkhorn_distance(nodes_g[: b // 2], nodes_g[b // 2:], eps, max_iter, red)[0]
    if return_debug:
        return losses, {"hidden": cur, "edges": edges, "nodes_g": nodes_g, "pred": of}
    return losses
