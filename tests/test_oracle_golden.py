"""CPU: the oracle restatements reproduce the outputs of the unmodified reference
(tests/golden/*.pt, written by oracle/make_golden.py).  Tolerances are fp32 round-off of
re-associated sums; integer outputs (k-NN indices, sampled labels) are bit-exact."""
import torch
import torch.nn.functional as F

from graphecho_b200 import synth
from oracle import graph_ops as G, vig_ops as V, fpn_ops as FP, gmodule_ops as GM, tgcn_ops as T
from oracle.params import make_params


def close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_affinity_and_sinkhorn(golden):
    g = golden("affinity_sinkhorn")
    p = make_params("affinity", fill_prefix="node_affinity.", requires_grad=True)
    X, Y = g["X"].clone().requires_grad_(), g["Y"].clone().requires_grad_()
    M = G.affinity(X, Y, p)
    close(M, g["M"])
    P = G.sinkhorn_rpm_exp(M, 20, True)
    close(P, g["P"], rtol=2e-4, atol=1e-6)
    (P * g["W"]).sum().backward()
    close(X.grad, g["dX"], rtol=1e-3, atol=1e-6)
    close(Y.grad, g["dY"], rtol=1e-3, atol=1e-6)
    close(p["fc_M.2.weight"].grad, g["dw2"], rtol=1e-3, atol=1e-6)
    close(p["fc_M.2.bias"].grad, g["db2"], rtol=1e-3, atol=1e-6)
    close(p["fc_M.0.weight"].grad[:8], g["dfc0"], rtol=1e-3, atol=1e-6)
    close(p["project_sr.weight"].grad[:8], g["dPs"], rtol=1e-3, atol=1e-6)
    close(G.sinkhorn_rpm(g["Mraw"], 5, True), g["rpm5"])
    close(G.sinkhorn_rpm(g["Mraw"], 3, False), g["rpm3_noslack"])


def test_forward_aff_and_qu(golden):
    g = golden("forward_aff")
    p = make_params("gmodule_nc3")
    a, b = g["n1"].clone().requires_grad_(), g["n2"].clone().requires_grad_()
    loss, Mn = G.forward_aff(a, b, g["l1"], g["l2"], p, 3)
    qu = G.forward_qu(g["e1"], g["e2"], Mn)
    close(Mn, g["Mn"], rtol=2e-4, atol=1e-6)
    close(loss, g["loss"])
    close(qu, g["qu"])
    (loss + qu).backward()
    close(a.grad, g["dn1"], rtol=1e-3, atol=1e-7)
    close(b.grad, g["dn2"], rtol=1e-3, atol=1e-7)


def test_attention(golden):
    g = golden("attention")
    p = make_params("mha", fill_prefix="intra_domain_graph.")
    out, att = G.mha_v2(g["key"], g["key"], g["query"], p)
    close(out, g["out"])
    close(att, g["attn"], atol=1e-7)


def test_sinkhorn_distance(golden):
    cases = golden("sinkhorn_distance")
    for name, c in cases.items():
        x, y = c["x"].clone().requires_grad_(), c["y"].clone().requires_grad_()
        cost, pi, C, nits = G.sinkhorn_distance(x, y, c["eps"], c["max_iter"], c["reduction"])
        close(C, c["C"])
        close(pi, c["pi"], rtol=1e-3, atol=1e-7)
        close(cost, c["cost"], rtol=1e-3, atol=1e-6)
        if "dx" in c:
            cost.sum().backward()
            close(x.grad, c["dx"], rtol=2e-3, atol=1e-6)
            close(y.grad, c["dy"], rtol=2e-3, atol=1e-6)
        assert 1 <= nits <= c["max_iter"]


def test_knn_mrconv_grapher(golden):
    g = golden("vig")
    assert torch.equal(V.dense_dilated_knn(g["x"], g["y"], 5, 2), g["e_xy"])
    assert torch.equal(V.dense_dilated_knn(g["x"], None, 9, 1, g["rel"]), g["e_self"])
    assert torch.equal(V.dense_dilated_knn(g["x"], None, 9, 1), g["e_plain"])
    p = make_params("mrconv32_64", fill_prefix="grapher.gconv.", requires_grad=True)
    x, y = g["x"].clone().requires_grad_(), g["y"].clone().requires_grad_()
    o = V.mrconv(x, g["e_xy"], p, "nn.", y=y, norm=None, act="gelu")
    close(o, g["mr_out"])
    o.square().sum().backward()
    close(x.grad, g["mr_dx"], rtol=1e-3, atol=1e-5)
    close(y.grad, g["mr_dy"], rtol=1e-3, atol=1e-5)
    for r in (1, 2):
        c = g[f"grapher_r{r}"]
        p = make_params("grapher32", fill_prefix=f"grapher_r{r}.", requires_grad=True)
        xin = c["x"].clone().requires_grad_()
        out = V.grapher(xin, p, "", k=5, dilation=1, r=r, norm="batch", act="gelu", training=True)
        close(out, c["out"])
        out.square().mean().backward()
        close(xin.grad, c["dx"], rtol=1e-3, atol=1e-6)
        close(p["fc1.0.weight"].grad, c["dfc1"], rtol=1e-3, atol=1e-6)
        close(p["fc2.1.running_mean"], c["rm"])
        close(p["fc1.1.running_var"], c["rv"])


def test_fpn_and_discriminator(golden):
    g = golden("fpn")
    for bb, contract, nc, hw in (("resnet", "fpn_resnet_nc1", 1, 112), ("VGG16", "fpn_vgg16_nc3", 3, 64)):
        x = g[f"{bb}_x"]
        for mode in ("eval", "train"):
            rec = g[f"{bb}_{mode}"]
            p = make_params(contract, scale=0.7, requires_grad=(mode == "train"))
            xr = x.clone().requires_grad_(mode == "train")
            logits, feats = FP.fpn_forward(xr, p, "resnet" if bb == "resnet" else "vgg16", training=(mode == "train"))
            close(logits, rec["logits"], rtol=1e-3, atol=1e-4)
            close(feats[3], rec["p5"], rtol=1e-3, atol=1e-4)
            close(feats[2], rec["p4"], rtol=1e-3, atol=1e-4)
            close(feats[0][:, ::16, ::3, ::3], rec["p2_slice"], rtol=1e-3, atol=1e-4)
            close(feats[1][:, ::16, ::2, ::2], rec["p3_slice"], rtol=1e-3, atol=1e-4)
            if mode == "train":
                B = x.shape[0]
                mask = synth.disc_masks(B, nc, hw) if nc > 1 else synth.disc_masks(B, 2, hw)[:, 1:2]
                loss = FP.seg_loss(logits, mask)
                close(loss, rec["loss"])
                loss.backward()
                close(p["conv3.weight"].grad, rec["dconv3"], rtol=5e-3, atol=1e-5)
                close(p["gn1.weight"].grad, rec["dgn1"], rtol=5e-3, atol=1e-5)
                close(p["semantic_branch.weight"].grad[:4], rec["dsem"], rtol=5e-3, atol=1e-5)
                close(p["toplayer.weight"].grad[:4, :64], rec["dtop"], rtol=5e-3, atol=1e-5)
                close(xr.grad, rec["dx"], rtol=5e-3, atol=1e-5)
                key = "back_bone.bn1.running_mean" if bb == "resnet" else "back_bone.block_1.1.running_mean"
                close(p[key], rec["bn1_rm"])
    d = golden("discriminator")
    p = make_params("discriminator", fill_prefix="dis.", requires_grad=True)
    a, b = d["fs"].clone().requires_grad_(), d["ft"].clone().requires_grad_()
    loss = FP.discriminator_loss(a, b, p, 0.02)
    close(loss, d["loss"])
    loss.backward()
    close(a.grad, d["dfs"], rtol=1e-3, atol=1e-8)
    close(b.grad, d["dft"], rtol=1e-3, atol=1e-8)
    close(p["cls_logits.weight"].grad, d["dcls"], rtol=1e-3, atol=1e-6)


def test_gmodule_and_sampler(golden):
    g = golden("gmodule")
    B, hw, nc = g["B"], g["hw"], g["nc"]
    fs = [f.requires_grad_() for f in synth.pyramid(B, hw, seed=21)]
    ft = [f.requires_grad_() for f in synth.pyramid(B, hw, seed=22)]
    masks, score = synth.disc_masks(B, nc, hw), synth.disc_masks(B, nc, hw, shift=6)
    p = make_params("gmodule_nc3", requires_grad=True)
    (n1, n2), losses = GM.gmodule_train(fs, ft, masks, score, p, nc, dropout=0.0, training=True)
    assert set(losses) == set(g["losses"])
    for k in losses:
        close(losses[k], g["losses"][k], rtol=1e-3, atol=1e-6)
    close(n1, g["n1"], rtol=1e-3, atol=1e-4)
    close(n2, g["n2"], rtol=1e-3, atol=1e-4)
    close(p["sr_seed"], g["sr_seed"], rtol=1e-3, atol=1e-4)
    close(p["tg_seed"], g["tg_seed"], rtol=1e-3, atol=1e-4)
    sum(losses.values()).backward()
    close(fs[3].grad, g["dfs3"], rtol=5e-3, atol=1e-7)
    close(fs[0].grad.abs().sum(), g["dfs0_abs"], rtol=5e-3, atol=1e-7)
    close(p["node_affinity.fc_M.2.weight"].grad, g["daff"], rtol=5e-3, atol=1e-7)
    s = golden("sampler")
    feats = synth.pyramid(B, hw, seed=21)
    nodes, labels, weights = GM.sample_nodes(GM.compute_locations(feats), feats, GM.find_bbox(masks), nc)
    assert len(labels) == s["count"]
    assert torch.equal(labels, s["labels"])
    close(nodes.sum(1), s["nodes_sum"])
    assert torch.equal(GM.find_bbox(masks)[1], s["boxes"])


def test_tgcn(golden):
    g = golden("tgcn")
    for transport, contract in (("node_discriminate", "tgcn_nd"), ("sinkhorn_distance", "tgcn_sd")):
        rec = g[transport]
        p = make_params(contract, requires_grad=True)
        feats = [f.requires_grad_() for f in synth.clip_pyramid(2, 3, 256, seed=31)]
        losses = T.tgcn_forward(feats, (rec["src"], rec["tgt"]), p, transport=transport, sinkhorn=(0.1, 5, "mean"),
                                training=True, dropout=0.0)
        assert set(losses) == set(rec["losses"])
        for k in losses:
            close(losses[k], rec["losses"][k], rtol=2e-3, atol=1e-6)
        sum(losses.values()).backward()
        close(feats[3].grad, rec["df3"], rtol=1e-2, atol=1e-8)
        close(feats[0].grad.abs().sum(), rec["df0_abs"], rtol=1e-2, atol=1e-8)
        close(p["pos_embed"].grad[:, :, :8], rec["dpos"], rtol=1e-2, atol=1e-8)
        close(p["grapher.MLP.1.running_mean"], rec["mlp_rm"])
        close(p["prediction.1.running_var"], rec["pred_rv"])
