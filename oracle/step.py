"""CPU restatement of one GraphEcho UDA training step on the oracle operators — the reported CPU
baseline (`bench.py` cpu_baseline / `--impl reference`) and the step-level parity checker.
Follows train_cardiac_uda.py:223-325 / train_camus_echo.py:206-299 with the workload definition of
graphecho_b200.engine (config 2 of BASELINE.json: FPN(resnet) + ViG Grapher on p2 + GModule + four
Discriminators on 112x112 clip frames).  Test / measurement infrastructure, never the product."""
from __future__ import annotations

import torch

from . import fpn_ops as FP, gmodule_ops as GM, vig_ops as V
from .params import make_params


def build_params(num_classes=2, backbone="resnet", grapher=True, seed_scale=1.0):
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
    return P


def leaves(pdict):
    return [t for t in pdict.values() if t.requires_grad]


def build_optimizers(P, lr_net=3e-4, lr_aux=2.5e-3, wd=1e-4):
    opt = {"fpn": torch.optim.Adam(leaves(P["fpn"]), lr=lr_net, betas=(0.9, 0.999), weight_decay=wd)}
    for k in P:
        if k != "fpn":
            opt[k] = torch.optim.SGD(leaves(P[k]), lr=lr_aux, momentum=0.9, weight_decay=wd)
    return opt


def forward_losses(P, frames_src, masks_src, frames_tgt, num_classes=2, backbone="resnet", dropout=0.1,
                   cluster=True, seg_weight=1.0):
    ns = frames_src.shape[0]
    logits, feats = FP.fpn_forward(torch.cat([frames_src, frames_tgt]), P["fpn"],
                                   "resnet" if backbone == "resnet" else "vgg16", training=True)
    losses = {"seg_loss": seg_weight * FP.seg_loss(logits[:ns], masks_src)}
    if "grapher" in P:
        feats = [V.grapher(feats[0], P["grapher"], "", k=9, dilation=1, r=1, norm="batch", act="gelu", training=True)] \
            + list(feats[1:])
    fs, ft = [f[:ns] for f in feats], [f[ns:] for f in feats]
    score = torch.where(torch.sigmoid(logits[ns:]) > 0.5, 1, 0)
    _, mid = GM.gmodule_train(fs, ft, masks_src, score, P["gm"], num_classes, dropout=dropout, training=True,
                              cluster=cluster)
    losses.update(mid)
    for i, lvl in enumerate(("p2", "p3", "p4", "p5")):
        losses[f"loss_adv_{lvl}"] = 0.1 * FP.discriminator_loss(fs[i], ft[i], P[f"dis_{lvl}"], 0.02)
    return losses


def train_step(P, opt, frames_src, masks_src, frames_tgt, **kw):
    for o in opt.values():
        o.zero_grad(set_to_none=True)
    losses = forward_losses(P, frames_src, masks_src, frames_tgt, **kw)
    total = sum(losses.values())
    total.backward()
    for o in opt.values():
        o.step()
    return total.detach(), {k: v.detach() for k, v in losses.items()}
