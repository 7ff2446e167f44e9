"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on seeded
inputs with name-keyed deterministic weights (oracle/detfill.py).

Run here (the build container) only:   python -m oracle.make_golden
The GPU box has no /root/reference; tests read the committed fixtures.  Test infrastructure.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("GRAPHECHO_REFERENCE", "/root/reference"))
OUT = ROOT / "tests" / "golden"

from oracle.detfill import fill_module  # noqa: E402
from graphecho_b200 import synth  # noqa: E402


def _import_reference():
    np.float = float  # vig.py:74 uses the removed alias
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(Path(__file__).resolve().parent / "refshim"))
    with contextlib.redirect_stdout(io.StringIO()):
        import models.fpnseg as fpnseg
        import models.graph_matching as gm
        import models.affinity_layer as aff
        import models.transformer as tr
        import models.vig as vig
        import models.TGCN as tgcn
        import utils.sinkhorn_distance as sd
        import utils.losses as losses
    return dict(fpnseg=fpnseg, gm=gm, aff=aff, tr=tr, vig=vig, tgcn=tgcn, sd=sd, losses=losses)


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _no_dropout(mod):
    for m in mod.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return mod


def save(name, obj):
    OUT.mkdir(parents=True, exist_ok=True)
    torch.save(obj, OUT / f"{name}.pt")
    size = (OUT / f"{name}.pt").stat().st_size
    print(f"  {name}.pt  {size / 1024:.1f} KiB")


def gen_affinity_sinkhorn(R):
    torch.manual_seed(11)
    X, Y = torch.randn(37, 256), torch.randn(45, 256)
    A = fill_module(R["aff"].Affinity(256), prefix="node_affinity.")
    Xr, Yr = X.clone().requires_grad_(), Y.clone().requires_grad_()
    M = A(Xr, Yr)
    gm = _quiet(R["gm"].GModule, 256, 3, "cpu")
    z = gm.InstNorm_layer(M[None, None])
    P = gm.sinkhorn_rpm(z[:, 0], n_iters=20).squeeze().exp()
    Wt = torch.randn(37, 45, generator=torch.Generator().manual_seed(5))
    (P * Wt).sum().backward()
    Mraw = torch.randn(1, 20, 31, generator=torch.Generator().manual_seed(6)) * 2
    plain5 = gm.sinkhorn_rpm(Mraw, n_iters=5)
    noslack = gm.sinkhorn_rpm(Mraw, n_iters=3, slack=False)
    save("affinity_sinkhorn", dict(X=X, Y=Y, M=M.detach(), P=P.detach(), W=Wt, dX=Xr.grad, dY=Yr.grad,
                                  dfc0=A.fc_M[0].weight.grad[:8].clone(), dw2=A.fc_M[2].weight.grad.clone(),
                                  db2=A.fc_M[2].bias.grad.clone(), dPs=A.project_sr.weight.grad[:8].clone(),
                                  Mraw=Mraw, rpm5=plain5, rpm3_noslack=noslack))


def gen_forward_aff(R):
    gm = fill_module(_quiet(R["gm"].GModule, 256, 3, "cpu"))
    torch.manual_seed(12)
    n1, n2 = torch.randn(40, 256), torch.randn(52, 256)
    l1 = torch.randint(0, 3, (40,)).float().sort()[0]
    l2 = torch.randint(0, 3, (52,)).float().sort()[0]
    a, b = n1.clone().requires_grad_(), n2.clone().requires_grad_()
    loss, Mn = gm._forward_aff(a, b, l1, l2)
    e1 = torch.softmax(torch.randn(40, 40), -1)
    e2 = torch.softmax(torch.randn(52, 52), -1)
    qu = gm._forward_qu(e1, e2, Mn)
    (loss + qu).backward()
    save("forward_aff", dict(n1=n1, n2=n2, l1=l1, l2=l2, e1=e1, e2=e2, loss=loss.detach(), Mn=Mn.detach(),
                             qu=qu.detach(), dn1=a.grad, dn2=b.grad))


def gen_attention(R):
    att = fill_module(R["tr"].MultiHeadAttention(256, 1, dropout=0.1, version="v2"), prefix="intra_domain_graph.").eval()
    torch.manual_seed(13)
    k, q = torch.randn(33, 256), torch.randn(21, 256)
    out, a = att(k, k, q)
    save("attention", dict(key=k, query=q, out=out.detach(), attn=a.detach()))


def gen_sinkhorn_distance(R):
    torch.manual_seed(14)
    cases = {}
    for name, (B, P1, P2, D, scale, eps, iters, red) in {
        "small": (3, 16, 12, 32, 0.15, 0.1, 5, "mean"),
        "tgcn": (2, 64, 64, 256, 0.05, 0.1, 5, "mean"),
        "long": (2, 10, 10, 8, 0.3, 0.1, 50, "none"),
    }.items():
        x, y = torch.randn(B, P1, D) * scale, torch.randn(B, P2, D) * scale
        xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
        cost, pi, C = R["sd"].SinkhornDistance(eps, iters, red)(xr, yr)
        cost.sum().backward()
        cases[name] = dict(x=x, y=y, eps=eps, max_iter=iters, reduction=red, cost=cost.detach(), pi=pi.detach(),
                           C=C.detach(), dx=xr.grad, dy=yr.grad)
    x2, y2 = torch.randn(9, 16) * 0.2, torch.randn(7, 16) * 0.2
    cost, pi, C = R["sd"].SinkhornDistance(0.1, 5, "none")(x2, y2)
    cases["2d"] = dict(x=x2, y=y2, eps=0.1, max_iter=5, reduction="none", cost=cost, pi=pi, C=C)
    save("sinkhorn_distance", cases)


def gen_vig(R):
    vig = R["vig"]
    torch.manual_seed(15)
    x, y = torch.randn(2, 32, 50, 1), torch.randn(2, 32, 40, 1)
    e_xy = vig.DenseDilatedKnnGraph(5, 2)(x, y)
    rel = torch.randn(1, 50, 50) * 0.05
    e_self = vig.DenseDilatedKnnGraph(9, 1)(x, None, rel)
    e_plain = vig.DenseDilatedKnnGraph(9, 1)(x)
    mr = fill_module(vig.MRConv2d(32, 64, "gelu", None, True), prefix="grapher.gconv.")
    xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
    o = mr(xr, e_xy, yr)
    o.square().sum().backward()
    out = dict(x=x, y=y, rel=rel, e_xy=e_xy, e_self=e_self, e_plain=e_plain, mr_out=o.detach(), mr_dx=xr.grad, mr_dy=yr.grad)
    for r in (1, 2):
        g = fill_module(vig.Grapher(32, 5, 1, "mr", "gelu", "batch", True, False, 0.0, r, 64, 0.0, False), prefix=f"grapher_r{r}.")
        g.train()
        xin = torch.randn(2, 32, 8, 8, generator=torch.Generator().manual_seed(150 + r))
        xg = xin.clone().requires_grad_()
        og = _quiet(g, xg)
        og.square().mean().backward()
        out[f"grapher_r{r}"] = dict(x=xin, out=og.detach(), dx=xg.grad,
                                    rm=g.fc2[1].running_mean.clone(), rv=g.fc1[1].running_var.clone(),
                                    dfc1=g.fc1[0].weight.grad.clone())
    save("vig", out)


def gen_fpn(R):
    fp = R["fpnseg"]
    out = {}
    for bb, nc, hw, B in (("resnet", 1, 112, 2), ("VGG16", 3, 64, 2)):
        net = fill_module(fp.FPN([2, 4, 23, 3], nc, 1, back_bone=bb), scale=0.7)
        x = torch.rand(B, 1, hw, hw, generator=torch.Generator().manual_seed(16))
        for mode in ("eval", "train"):
            net.train(mode == "train")
            if mode == "train":
                fill_module(net, scale=0.7)   # reset running stats moved by nothing yet (idempotent)
            xr = x.clone().requires_grad_(mode == "train")
            logits, feats = net(xr)
            rec = dict(logits=logits.detach().clone(), p5=feats[3].detach().clone(),
                       p4=feats[2].detach().clone(), p2_slice=feats[0][:, ::16, ::3, ::3].detach().clone(),
                       p3_slice=feats[1][:, ::16, ::2, ::2].detach().clone())
            if mode == "train":
                mask = synth.disc_masks(B, nc, hw) if nc > 1 else synth.disc_masks(B, 2, hw)[:, 1:2]
                loss = R["losses"].DiceLoss()(logits, mask) + torch.nn.functional.binary_cross_entropy_with_logits(logits, mask)
                loss.backward()
                rec.update(loss=loss.detach(), dx=xr.grad.clone(), dconv3=net.conv3.weight.grad.clone(),
                           dgn1=net.gn1.weight.grad.clone(), dsem=net.semantic_branch.weight.grad[:4].clone(),
                           dtop=net.toplayer.weight.grad[:4, :64].clone(),
                           bn1_rm=(net.back_bone.bn1.running_mean if bb == "resnet" else net.back_bone.block_1[1].running_mean).clone())
            out[f"{bb}_{mode}"] = rec
        out[f"{bb}_x"] = x
    save("fpn", out)
    d = fill_module(fp.Discriminator(grad_reverse_lambda=0.02), prefix="dis.")
    torch.manual_seed(17)
    fs, ft = torch.randn(2, 256, 8, 8), torch.randn(2, 256, 8, 8)
    a, b = fs.clone().requires_grad_(), ft.clone().requires_grad_()
    loss = d((a, b))
    loss.backward()
    save("discriminator", dict(fs=fs, ft=ft, loss=loss.detach(), dfs=a.grad, dft=b.grad,
                               dcls=d.cls_logits.weight.grad.clone()))


def gen_gmodule(R):
    nc = 3
    gm = fill_module(_no_dropout(_quiet(R["gm"].GModule, 256, nc, "cpu")))
    gm.train()
    B, hw = 2, 256
    feats_s = synth.pyramid(B, hw, seed=21)
    feats_t = synth.pyramid(B, hw, seed=22)
    masks = synth.disc_masks(B, nc, hw)
    score = synth.disc_masks(B, nc, hw, shift=6)
    fs = [f.clone().requires_grad_() for f in feats_s]
    ft = [f.clone().requires_grad_() for f in feats_t]
    _, (n1, n2), losses = gm(None, (fs, ft), targets=masks, score_maps=score)
    sum(losses.values()).backward()
    save("gmodule", dict(B=B, hw=hw, nc=nc, losses={k: v.detach() for k, v in losses.items()},
                         n1=n1.detach(), n2=n2.detach(), sr_seed=gm.sr_seed.clone(), tg_seed=gm.tg_seed.clone(),
                         dfs3=fs[3].grad.clone(), dft0_sum=ft[0].grad.sum(), dfs0_abs=fs[0].grad.abs().sum(),
                         daff=gm.node_affinity.fc_M[2].weight.grad.clone()))
    # sampler-only record (labels are integer work: bit-exact)
    locs = gm.compute_locations(feats_s)
    nodes, labels, weights = gm.graph_generator(locs, feats_s, gm.find_bbox(masks))
    save("sampler", dict(labels=labels, nodes_sum=nodes.sum(1), count=len(labels), boxes=gm.find_bbox(masks)[1]))


def gen_tgcn(R):
    out = {}
    for transport in ("node_discriminate", "sinkhorn_distance"):
        m = fill_module(_no_dropout(_quiet(R["tgcn"].TGCN, 256, 256, (3, 8, 8), 10, 10, None, transport)))
        m.train()
        feats = synth.clip_pyramid(2, 3, 256, seed=31)
        fr = [f.clone().requires_grad_() for f in feats]
        torch.manual_seed(32)
        src, tgt = torch.randn(17, 256), torch.randn(23, 256)
        sink = R["sd"].SinkhornDistance(0.1, 5, "mean")
        idx = (torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))
        losses = m(fr, (src, tgt), sink, torch.nn.CrossEntropyLoss(), idx, r=[8, 4, 2, 1])
        sum(losses.values()).backward()
        out[transport] = dict(losses={k: v.detach() for k, v in losses.items()}, src=src, tgt=tgt,
                              df3=fr[3].grad.clone(), df0_abs=fr[0].grad.abs().sum(),
                              dpos=m.pos_embed.grad[:, :, :8].clone(),
                              mlp_rm=m.grapher.MLP[1].running_mean.clone(), pred_rv=m.prediction[1].running_var.clone())
    save("tgcn", out)


def stage_probe(x):
    """A small deterministic sub-sample of a [B,C,H,W] activation (<= 2x8x4x4 values) + its mean / std."""
    B, C, H, W = x.shape
    sub = x[:, ::max(1, C // 8), ::max(1, H // 4), ::max(1, W // 4)][:, :8, :4, :4].contiguous()
    return dict(sub=sub.clone(), mean=x.mean().clone(), std=x.std().clone())


def gen_pvig(R):
    """pvig_ti_224_gelu (DeepGCN), eval mode, seeded 224x224 input.  Two kinds of records:
    * chained: a probe of the activation after the stem (+pos_embed) and after every backbone element, and the logits;
    * teacher-forced, for the elements of stages 3-4 (dilated k-NN: the strided pick over the distance-sorted list makes
      the selected set sensitive to ANY adjacent near-tie, so one fp32 round-off flip early on is amplified block by
      block in a chained run): the reference element's input (image 0, rounded to fp16) and a probe of ITS output for
      exactly that input; plus the pooled feature vector and the logits of the prediction head."""
    net = fill_module(_quiet(R["vig"].pvig_ti_224_gelu), scale=0.7).eval()
    x = torch.rand(2, 3, 224, 224, generator=torch.Generator().manual_seed(41))
    stages, forced = [], {}
    with torch.no_grad():
        h = net.stem(x) + net.pos_embed
        stages.append(stage_probe(h))
        for i, blk in enumerate(net.backbone):
            if i >= 6:
                inp = h[:1].half()
                forced[i] = dict(inp=inp, out=stage_probe(_quiet(blk, inp.float())))
            h = _quiet(blk, h)
            stages.append(stage_probe(h))
        pooled = torch.nn.functional.adaptive_avg_pool2d(h, 1)
        logits = _quiet(net, x)
        head = net.prediction(pooled).squeeze(-1).squeeze(-1)
    save("pvig", dict(seed=41, shape=[2, 3, 224, 224], logits=logits, stages=stages, forced=forced,
                      pooled=pooled, head_logits=head))   # x = torch.rand(shape, Generator(seed))


def gen_state_contract(R):
    """state_dict keys / shapes of every hot-path module, as the reference builds them."""
    import json
    vig = R["vig"]
    mods = {
        "fpn_resnet_nc1": R["fpnseg"].FPN([2, 4, 23, 3], 1, 1, back_bone="resnet"),
        "fpn_vgg16_nc3": R["fpnseg"].FPN([2, 4, 23, 3], 3, 1, back_bone="VGG16"),
        "discriminator": R["fpnseg"].Discriminator(grad_reverse_lambda=0.02),
        "gmodule_nc3": _quiet(R["gm"].GModule, 256, 3, "cpu"),
        "tgcn_nd": _quiet(R["tgcn"].TGCN, 256, 256, (3, 8, 8), 10, 10, None, "node_discriminate"),
        "tgcn_sd": _quiet(R["tgcn"].TGCN, 256, 256, (3, 8, 8), 10, 10, None, "sinkhorn_distance"),
        "grapher32": vig.Grapher(32, 5, 1, "mr", "gelu", "batch", True, False, 0.0, 1, 64, 0.0, False),
        "grapher256": vig.Grapher(256, 9, 1, "mr", "gelu", "batch", True, False, 0.0, 1, 784, 0.0, False),
        "mrconv32_64": vig.MRConv2d(32, 64, "gelu", None, True),
        "affinity": R["aff"].Affinity(256),
        "mha": R["tr"].MultiHeadAttention(256, 1, dropout=0.1, version="v2"),
    }
    contract = {k: {n: list(t.shape) for n, t in m.state_dict().items()} for k, m in mods.items()}
    contract["_param_counts"] = {k: sum(p.numel() for p in m.parameters()) for k, m in mods.items()}
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "state_contract.json").write_text(json.dumps(contract, indent=0))
    print("  state_contract.json", {k: v for k, v in contract["_param_counts"].items()})


def main():
    R = _import_reference()
    print("writing fixtures to", OUT)
    gen_state_contract(R)
    gen_affinity_sinkhorn(R)
    gen_forward_aff(R)
    gen_attention(R)
    gen_sinkhorn_distance(R)
    gen_vig(R)
    gen_fpn(R)
    gen_gmodule(R)
    gen_tgcn(R)
    gen_pvig(R)


if __name__ == "__main__":
    main()
