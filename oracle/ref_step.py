"""One UDA training step on the reference's OWN modules (imported unmodified from oracle/_ref, staged by
oracle/build_ref.py, or from /root/reference when that is mounted), composed in the order of the reference
trainer's inner loop (train_cardiac_uda.py:223-325; the trainers themselves cannot run: monai / tensorboardX /
data are absent).  Two uses:
  * `bench.py --impl reference` / cpu_baseline.kind == "reference": the reference implementation timed on the host;
  * tests/test_ref_step.py: pins oracle/step.py (the restatement) to the reference step, loss for loss.
Weights are the same name-keyed deterministic fill (oracle/detfill.py) every other side uses.
Test / measurement infrastructure, never the product."""
from __future__ import annotations

import contextlib
import io
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from .build_ref import OUT as REF_STAGED, available as staged_available
from .detfill import fill_module

_R = None


def reference_root() -> Path | None:
    if staged_available():
        return REF_STAGED
    live = Path(os.environ.get("GRAPHECHO_REFERENCE", "/root/reference"))
    return live if (live / "models" / "fpnseg.py").exists() else None


def load():
    """Import the reference modules (namespace packages `models`, `utils` of the reference tree + the timm stub)."""
    global _R
    if _R is not None:
        return _R
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is not available (no oracle/_ref and no /root/reference)")
    if not hasattr(np, "float"):
        np.float = float                    # vig.py:74 uses the alias numpy removed
    for p in (str(root), str(Path(__file__).resolve().parent / "refshim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):
        import models.fpnseg as fpnseg
        import models.graph_matching as gm
        import models.vig as vig
        import models.TGCN as tgcn
        import utils.sinkhorn_distance as sd
        import utils.losses as losses
    for mod in (fpnseg, gm, vig, tgcn, sd, losses):
        if not str(Path(mod.__file__).resolve()).startswith(str(root.resolve())):
            raise RuntimeError(f"{mod.__name__} resolved to {mod.__file__}, not to the reference tree {root}")
    _R = SimpleNamespace(fpnseg=fpnseg, gm=gm, vig=vig, tgcn=tgcn, sd=sd, losses=losses, root=root)
    return _R


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def build_modules(num_classes=2, backbone="resnet", grapher=True, tgcn=False, clip_frames=8, hw=112, dropout=None):
    R = load()
    M = {"net": fill_module(R.fpnseg.FPN([2, 4, 23, 3], num_classes=num_classes, in_channel=1, back_bone=backbone), scale=0.7),
         "gm": fill_module(_quiet(R.gm.GModule, 256, num_classes, "cpu"))}
    for lvl in ("p2", "p3", "p4", "p5"):
        M[f"dis_{lvl}"] = fill_module(R.fpnseg.Discriminator(grad_reverse_lambda=0.02), prefix=f"dis_{lvl}.")
    if grapher:
        M["grapher"] = fill_module(R.vig.Grapher(256, 9, 1, "mr", "gelu", "batch", True, False, 0.0, 1, (hw // 4) ** 2,
                                                 0.0, False), prefix="grapher.")
    if tgcn:
        M["tgcn"] = fill_module(_quiet(R.tgcn.TGCN, 256, 256, (clip_frames, 8, 8), 10, 10), prefix="tgcn.")
    for m in M.values():
        m.train()
        if dropout is not None:
            for s in m.modules():
                if isinstance(s, torch.nn.Dropout):
                    s.p = dropout
    M["_loss"] = SimpleNamespace(dice=R.losses.DiceLoss(), bce=torch.nn.BCEWithLogitsLoss(reduction="mean"),
                                 ce=torch.nn.CrossEntropyLoss(), sinkhorn=R.sd.SinkhornDistance(eps=0.1, max_iter=5, reduction="mean"))
    return M


def build_optimizers(M, lr_net=3e-4, lr_aux=2.5e-3, wd=1e-4):
    opt = {"net": torch.optim.Adam(M["net"].parameters(), lr=lr_net, betas=(0.9, 0.999), weight_decay=wd)}
    for k, m in M.items():
        if k not in ("net", "_loss"):
            opt[k] = torch.optim.SGD(m.parameters(), lr=lr_aux, momentum=0.9, weight_decay=wd)
    return opt


def _network(M, x):
    pred, feats = M["net"](x)
    if "grapher" in M:                                         # config 2: the ViG Grapher on p2, part of the network call
        feats = [_quiet(M["grapher"], feats[0])] + list(feats[1:])
    return pred, feats


def forward_losses(M, frames_src, masks_src, frames_tgt, seg_weight=1.0, sinkhorn_nodes=False, sinkhorn_weight=0.001,
                   temporal=None):
    L = M["_loss"]
    losses = {}
    pred_source, features_source = _network(M, frames_src)                                       # :225
    losses["seg_loss"] = seg_weight * (L.dice(pred_source, masks_src) + L.bce(pred_source, masks_src))   # :228
    pred_target, features_target = _network(M, frames_tgt)                                       # :234
    score_maps = torch.where(torch.nn.Sigmoid()(pred_target) > 0.5, 1, 0)                        # :235
    (features_s, features_t), nodes, middle = _quiet(
        M["gm"], (frames_src, frames_tgt), (features_source, features_target), targets=masks_src, score_maps=score_maps)
    if sinkhorn_nodes and nodes[0].dim() == 2 and nodes[0].size(0) >= 6 and nodes[1].size(0) > 0:
        middle["sinkhorn_loss"] = sinkhorn_weight * L.sinkhorn(nodes[0], nodes[1])[0]
    losses.update(middle)
    for layer, name in enumerate(["p2", "p3", "p4", "p5"]):                                      # :241-243
        losses[f"loss_adv_{name}"] = 0.1 * M[f"dis_{name}"]((features_s[layer], features_t[layer]))
    if temporal is not None:                                                                     # :258-311
        imgs_temp, source_temp_masks, (b, t) = temporal
        nst = source_temp_masks.shape[0]
        masks_select = torch.where(torch.sum(source_temp_masks, dim=(1, 2, 3)) > 100, 1, 0)
        preds_, features_ = _network(M, imgs_temp)
        pred_source_temp = preds_[:nst]
        source_masks_ = torch.cat([(source_temp_masks[i] if ok else pred_source_temp[i]).unsqueeze(0)
                                   for i, ok in enumerate(masks_select)], dim=0)
        sf = [f[: f.shape[0] // 2] for f in features_]
        tf = [f[f.shape[0] // 2:] for f in features_]
        (_, _), (source_nodes, target_nodes), temp_middle = _quiet(
            M["gm"], (imgs_temp[:nst], imgs_temp[nst:]), (sf, tf), targets=source_masks_, score_maps=preds_[nst:])
        graph_features = [f.reshape(b, -1, *f.shape[1:]) for f in features_]
        tl = _quiet(M["tgcn"], graph_features, (source_nodes.clone().detach(), target_nodes.clone().detach()),
                    L.sinkhorn, L.ce, (None, None), r=[8, 4, 2, 1])
        losses["temporal_graph_loss"] = sum(tl.values()) + sum(temp_middle.values())
    return losses


def train_step(M, opt, frames_src, masks_src, frames_tgt, **kw):
    for o in opt.values():
        o.zero_grad()
    losses = forward_losses(M, frames_src, masks_src, frames_tgt, **kw)
    total = sum(losses.values())
    total.backward()                                                                             # :319
    for o in opt.values():
        o.step()
    return total.detach(), {k: v.detach() for k, v in losses.items()}
