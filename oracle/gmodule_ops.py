"""Oracle restatement of GModule._forward_train and its node sampler (rows a5-a8, a14 of
SURVEY.md §8), functional over a reference-keyed state dict.  Reproduces the reference's quirks
(location strides 8..128 on a stride-4..32 pyramid, bbox-based labels, float linspace picks).
Plain fp32 PyTorch + sklearn (update_seed), as the reference.  Test infrastructure."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import graph_ops as G
from .fpn_ops import grad_reverse

INF = 100000000
LOC_STRIDES = (8, 16, 32, 64, 128)                    # graph_matching.py:611
SIZES_OF_INTEREST = ((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF))   # :875-881


def compute_locations(features):
    """graph_matching.py:609-635: per level [h*w, 2] (x, y) = index*stride + stride//2."""
    out = []
    for lvl, f in enumerate(features):
        h, w = f.shape[-2:]
        s = LOC_STRIDES[lvl]
        ys = torch.arange(0, h * s, step=s, dtype=torch.float32, device=f.device)
        xs = torch.arange(0, w * s, step=s, dtype=torch.float32, device=f.device)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        out.append(torch.stack((gx.reshape(-1), gy.reshape(-1)), dim=1) + s // 2)
    return out


def masks_to_boxes(masks):
    """graph_matching.py:702-740: [K,H,W] -> [K,4] (xmin,ymin,xmax,ymax); empty mask -> (0,0,W,H)."""
    k, h, w = masks.shape
    boxes = torch.zeros((k, 4), dtype=torch.float, device=masks.device)
    for i in range(k):
        ys, xs = torch.where(masks[i] != 0)
        if xs.numel() == 0:
            boxes[i] = torch.tensor([0, 0, w, h], dtype=torch.float)
        else:
            boxes[i] = torch.stack([xs.min(), ys.min(), xs.max(), ys.max()]).float()
    return boxes


def find_bbox(masks):
    return [masks_to_boxes(m) for m in masks]                                     # :742-746


def location_labels(locations, boxes_per_image, num_class):
    """prepare_targets + compute_targets_for_locations (graph_matching.py:874-959): per level, the
    labels of every location of every image (image-major), smallest-area containing box wins."""
    pts = torch.cat(locations, dim=0)
    soi = torch.cat([pts.new_tensor(SIZES_OF_INTEREST[l])[None].expand(len(loc), -1)
                     for l, loc in enumerate(locations)], dim=0)
    xs, ys = pts[:, 0], pts[:, 1]
    per_image = []
    for boxes in boxes_per_image:
        area = torch.tensor([float((b[3] - b[1]) * (b[2] - b[0])) for b in boxes])          # :925-928
        reg = torch.stack([xs[:, None] - boxes[:, 0][None], ys[:, None] - boxes[:, 1][None],
                           boxes[:, 2][None] - xs[:, None], boxes[:, 3][None] - ys[:, None]], dim=2)
        inside = reg.min(dim=2)[0] > 0
        mx = reg.max(dim=2)[0]
        cared = (mx >= soi[:, [0]]) & (mx <= soi[:, [1]])
        a = area[None].repeat(len(pts), 1)
        a[inside == 0] = INF
        a[cared == 0] = INF
        amin, ainds = a.min(dim=1)
        lab = torch.arange(num_class)[ainds]
        lab[amin == INF] = 0
        per_image.append(lab)
    sizes = [len(l) for l in locations]
    split = [torch.split(lab, sizes, dim=0) for lab in per_image]
    return [torch.cat([s[l] for s in split], dim=0) for l in range(len(locations))]


def sample_nodes(locations, features, boxes_per_image, num_class, per_class=100, bg_ratio=8):
    """PrototypeComputation.__call__, `locations` branch (graph_matching.py:971-1013)."""
    labels = location_labels(locations, boxes_per_image, num_class)
    c = features[0].shape[1]
    pos_pts, pos_lab, neg_pts = [], [], []
    for l, lab in enumerate(labels):
        flat = features[l].permute(0, 2, 3, 1).reshape(-1, c)
        pos, neg = lab.reshape(-1) > 0, lab.reshape(-1) == 0
        pa, la = flat[pos], lab[pos]
        step = len(la) // per_class
        if step > 1:
            pa, la = pa[::step], la[::step]
        pos_pts.append(pa)
        pos_lab.append(la)
        npos = len(pa)
        nneg_all = int(neg.sum())
        cand = flat[neg]
        if int(pos.sum()) > nneg_all:
            neg_pts.append(cand)
        else:
            pick = np.floor(np.linspace(0, nneg_all - 2, npos // bg_ratio)).astype(np.int64)      # :1001
            neg_pts.append(cand[torch.as_tensor(pick, dtype=torch.long)])
    pos_pts, pos_lab, neg_pts = torch.cat(pos_pts), torch.cat(pos_lab), torch.cat(neg_pts)
    nodes = torch.cat([neg_pts, pos_pts], dim=0)
    lab = torch.cat([pos_lab.new_zeros(len(neg_pts)), pos_lab])
    return nodes, lab, torch.ones_like(lab).long()


def regroup_by_class(nodes, labels, weights, p, sr_seed, tg_seed, generator=None):
    """_forward_preprocessing_source_target (graph_matching.py:381-483): class-major regrouping and
    hallucination of classes missing in one domain from the seed bank."""
    (sn, tn), (sl, tl), (sw, tw) = nodes, labels, weights
    S, T, SL, TL, SW, TW = [], [], [], [], [], []

    def halluc(seed_row, other):
        n = len(other)
        base = seed_row[None].expand(n, 256)
        if n < 5:
            noise = torch.normal(0, 0.01, size=other.size(), generator=generator)
            out = noise.to(base.device) + base
        else:
            out = torch.normal(mean=base, std=other.std(0)[None].expand(base.size()), generator=generator)
        return F.linear(out, p["seed_project_left.weight"], p["seed_project_left.bias"])

    for c in torch.cat([sl, tl]).unique():
        si, ti = sl == c, tl == c
        s_c, t_c = sn[si], tn[ti]
        if si.any() and ti.any():
            S.append(s_c); T.append(t_c)
            SL.append(s_c.new_ones(len(s_c)) * c); TL.append(t_c.new_ones(len(t_c)) * c)
            SW.append(sw[si]); TW.append(tw[ti])
        elif ti.any():
            n = len(t_c)
            S.append(halluc(sr_seed[int(c.item())], t_c)); T.append(t_c)
            SL.append(torch.ones(n) * c); TL.append(torch.ones(n) * c)
            SW.append(torch.ones(n, dtype=torch.long)); TW.append(tw[ti])
        elif si.any():
            n = len(s_c)
            S.append(s_c); T.append(halluc(tg_seed[int(c.item())], s_c))
            SL.append(torch.ones(n) * c); TL.append(torch.ones(n) * c)
            SW.append(sw[si]); TW.append(torch.ones(n, dtype=torch.long))
    return (torch.cat(S), torch.cat(T)), (torch.cat(SL), torch.cat(TL)), (torch.cat(SW), torch.cat(TW))


def update_seed(sr_nodes, sr_labels, tg_nodes, tg_labels, sr_seed, tg_seed, cluster=True, k=20):
    """GModule.update_seed (graph_matching.py:532-567): per class, (spectrally filtered) mean with
    cosine-similarity momentum.  Mutates the seed banks in place, like the reference buffers."""
    import sklearn.cluster as skc

    def one(nodes, labels, bank):
        for cls in labels.unique().long():
            bs = nodes[labels == cls].detach()
            if len(bs) > k and cluster:
                sp = skc.SpectralClustering(2, affinity="nearest_neighbors", n_jobs=-1, assign_labels="kmeans",
                                            random_state=1234, n_neighbors=len(bs) // 2)
                idx = sp.fit_predict(torch.cat([bank[cls][None, :], bs]).cpu().numpy())
                keep = torch.as_tensor((idx == idx[0])[1:])
                bs = bs[keep].mean(0)
            else:
                bs = bs.mean(0)
            mom = F.cosine_similarity(bs[None], bank[cls][None])
            bank[cls] = bank[cls] * mom + bs * (1.0 - mom)

    one(sr_nodes, sr_labels, sr_seed)
    if tg_nodes is not None:
        one(tg_nodes, tg_labels, tg_seed)


def gmodule_train(features_s, features_t, targets, score_maps, p, num_classes, dropout=0.0,
                  training=True, cluster=True, generator=None, return_debug=False):
    """GModule._forward_train (graph_matching.py:244-352) with the default flags (:110-138):
    o2o matching, node discriminator at 'feat', complete graph, domain interaction, quadratic
    matching.  p holds the GModule state dict (incl. sr_seed / tg_seed buffers, updated in place).
    Returns ((nodes_1, nodes_2), loss_dict)."""
    losses = {}
    n1, l1, w1 = sample_nodes(compute_locations(features_s), features_s, find_bbox(targets), num_classes)
    n2, l2, w2 = sample_nodes(compute_locations(features_t), features_t, find_bbox(score_maps), num_classes)
    if n1.size(0) < 6 or n1.dim() == 1:
        return (n1, n2), losses
    both = grad_reverse(torch.cat([n1, n2], dim=0), 0.02)                        # lambda_dis :125, :263-270
    tgt = torch.cat([torch.ones(len(n1)), torch.zeros(len(n2))]).to(both.device)
    losses["dis_loss"] = 0.1 * F.binary_cross_entropy_with_logits(G.node_dis(both, p).view(-1), tgt)
    n1, n2 = G.head_in_ln(n1, p), G.head_in_ln(n2, p)                            # :284-285
    (n1, n2), (l1, l2), (w1, w2) = regroup_by_class((n1, n2), (l1, l2), (w1, w2), p, p["sr_seed"],
                                                    p["tg_seed"], generator)
    n1, e1 = G.mha_v2(n1, n1, n1, p, "intra_domain_graph.", dropout, training)   # :295-296
    n2, e2 = G.mha_v2(n2, n2, n2, p, "intra_domain_graph.", dropout, training)
    update_seed(n1, l1, n2, l2, p["sr_seed"], p["tg_seed"], cluster)              # :298
    n2x = G.mha_v2(n1, n1, n2, p, "cross_domain_graph.", dropout, training)[0]   # :500-501
    n1x = G.mha_v2(n2, n2, n1, p, "cross_domain_graph.", dropout, training)[0]
    n1, n2 = n1x, n2x
    logits = G.node_cls(torch.cat([n1, n2], dim=0), p)                            # :505-530 (weights given -> 'none' + mean)
    losses["node_loss"] = 1.0 * F.cross_entropy(logits, torch.cat([l1, l2]).long(), reduction="none").float().mean()
    aff_loss, Mn = G.forward_aff(n1, n2, l1, l2, p, num_classes)                  # :345-346
    losses["mat_loss_aff"] = 0.1 * aff_loss
    losses["mat_loss_qu"] = G.forward_qu(e1.detach(), e2.detach(), Mn)            # :349-350
    if return_debug:
        return (n1, n2), losses, {"labels": (l1, l2), "edges": (e1, e2), "M": Mn}
    return (n1, n2), losses
