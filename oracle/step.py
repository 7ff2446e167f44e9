"""CPU restatement of one GraphEcho UDA training step on the oracle operators — the step-level parity
checker and the `port` CPU baseline of bench.py.  Follows train_cardiac_uda.py:223-325 /
train_camus_echo.py:206-299 call for call: the network runs on the source batch and on the target batch in
TWO separate train-mode calls (:225, :234 -> per-domain BatchNorm statistics, two running-stat updates), then
graph matching, the four discriminators, and (temporal_graph) a third network call on the [source | target]
clips followed by a second graph-matching call and TGCN (:258-311).  Workload definitions = graphecho_b200.engine
presets (BASELINE.json configs 2-5).  Test / measurement infrastructure, never the product."""
from __future__ import annotations

import torch

from . import fpn_ops as FP, gmodule_ops as GM, graph_ops as G, tgcn_ops as TG, vig_ops as V
from .params import make_params


def build_params(num_classes=2, backbone="resnet", grapher=True, tgcn=False, clip_frames=8, seed_scale=1.0):
    nc = num_classes
    name = "fpn_resnet_nc1" if backbone == "resnet" else "fpn_vgg16_nc3"
    P = {"fpn": make_params(name, scale=0.7, requires_grad=True,
                            overrides={"conv3.weight": [nc, 128, 1, 1], "conv3.bias": [nc]}),
         "gm": make_params("gmodule_nc3", requires_grad=True,
                           overrides={"node_cls_middle.2.weight": [nc, 512], "node_cls_middle.2.bias": [nc],
                                      "sr_seed": [nc, 256], "tg_seed": [nc, 256]})}
    for lvl in ("p2", "p3", "p4", "p5"):
        P[f"dis_{lvl}"] = make_params("discriminator", fill_prefix=f"dis_{lvl}.", requires_grad=True)
    if grapher:
        P["grapher"] = make_params("grapher256", fill_prefix="grapher.", requires_grad=True)
    if tgcn:
        P["tgcn"] = make_params("tgcn_nd", fill_prefix="tgcn.", requires_grad=True,
                                overrides={"pos_embed": [clip_frames, 1, 256, 8, 8]})
    return P


def leaves(pdict):
    return [t for t in pdict.values() if t.requires_grad]


def build_optimizers(P, lr_net=3e-4, lr_aux=2.5e-3, wd=1e-4):
    opt = {"fpn": torch.optim.Adam(leaves(P["fpn"]), lr=lr_net, betas=(0.9, 0.999), weight_decay=wd)}
    for k in P:
        if k != "fpn":
            opt[k] = torch.optim.SGD(leaves(P[k]), lr=lr_aux, momentum=0.9, weight_decay=wd)
    return opt


def _network(x, P, backbone):
    """self.network(x) (+ the config-2 ViG Grapher on p2, which this build treats as part of the network call)."""
    logits, feats = FP.fpn_forward(x, P["fpn"], "resnet" if backbone == "resnet" else "vgg16", training=True)
    if "grapher" in P:
        feats = [V.grapher(feats[0], P["grapher"], "", k=9, dilation=1, r=1, norm="batch", act="gelu", training=True)] \
            + list(feats[1:])
    return logits, feats


def forward_losses(P, frames_src, masks_src, frames_tgt, num_classes=2, backbone="resnet", dropout=0.1,
                   cluster=True, seg_weight=1.0, sinkhorn_nodes=False, sinkhorn_weight=0.001, temporal=None,
                   return_debug=False):
    pred_s, fs = _network(frames_src, P, backbone)                                            # :225
    losses = {"seg_loss": seg_weight * FP.seg_loss(pred_s, masks_src)}                        # :228
    pred_t, ft = _network(frames_tgt, P, backbone)                                            # :234
    score = torch.where(torch.sigmoid(pred_t) > 0.5, 1, 0)                                    # :235
    nodes, mid = GM.gmodule_train(fs, ft, masks_src, score, P["gm"], num_classes, dropout=dropout, training=True,
                                  cluster=cluster)                                            # :237
    if sinkhorn_nodes and nodes[0].dim() == 2 and nodes[0].size(0) >= 6 and nodes[1].size(0) > 0:
        mid["sinkhorn_loss"] = sinkhorn_weight * G.sinkhorn_distance(nodes[0], nodes[1], 0.1, 5, "mean")[0]
    losses.update(mid)
    for i, lvl in enumerate(("p2", "p3", "p4", "p5")):                                        # :241-243
        losses[f"loss_adv_{lvl}"] = 0.1 * FP.discriminator_loss(fs[i], ft[i], P[f"dis_{lvl}"], 0.02)
    debug = {"nodes": nodes, "logits_s": pred_s, "logits_t": pred_t}
    if temporal is not None:                                                                  # :258-311
        frames_temp, masks_temp, (b, t) = temporal
        nst = masks_temp.shape[0]
        preds, feats = _network(frames_temp, P, backbone)                                     # :280 (one call)
        avail = (masks_temp.sum(dim=(1, 2, 3)) > 100).view(-1, 1, 1, 1)                        # :277, 283-290
        targets = torch.where(avail, masks_temp, preds[:nst].detach())
        fts, ftt = [f[:nst] for f in feats], [f[nst:] for f in feats]                         # :292-295
        (n1, n2), tmid = GM.gmodule_train(fts, ftt, targets, preds[nst:], P["gm"], num_classes, dropout=dropout,
                                          training=True, cluster=cluster)                     # :297-298 raw logits as score maps
        total = sum(tmid.values()) if tmid else None
        if n1.numel() > 0 and n1.dim() == 2:
            gf = [f.reshape(b, t, *f.shape[1:]) for f in feats]                               # :300-302
            tl = TG.tgcn_forward(gf, (n1.detach(), n2.detach()), P["tgcn"], rs=(8, 4, 2, 1), training=True,
                                 dropout=dropout)                                             # :304
            tsum = sum(tl.values())
            total = tsum if total is None else total + tsum
            debug["tgcn"] = tl
        if total is not None:
            losses["temporal_graph_loss"] = total                                             # :309-311
    return (losses, debug) if return_debug else losses


def train_step(P, opt, frames_src, masks_src, frames_tgt, **kw):
    for o in opt.values():
        o.zero_grad(set_to_none=True)
    losses = forward_losses(P, frames_src, masks_src, frames_tgt, **kw)
    total = sum(losses.values())
    total.backward()
    for o in opt.values():
        o.step()
    return total.detach(), {k: v.detach() for k, v in losses.items()}
