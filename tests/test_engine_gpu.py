"""GPU: the whole training step (graphecho_b200.engine.UDAEngine) — step-level parity with the CPU oracle step
(oracle/step.py, itself pinned to a step of the unmodified reference modules by tests/test_ref_step.py) for
BASELINE.json configs 2, 3 and 4: every loss and a set of parameter gradients; a bf16 training step against the fp32
oracle; a multi-step training trajectory against the oracle's; CUDA-graph replay vs eager execution (losses,
gradients, buffers); phased vs single-graph backward; optimizer progress."""
import contextlib
import io

import pytest
import torch

from graphecho_b200 import synth
from graphecho_b200.engine import (EngineConfig, UDAEngine, make_batch, make_frame_batch, preset, split_streams,
                                   temporal_input)
from oracle import step as OS
from oracle.detfill import fill_module

pytestmark = pytest.mark.gpu


def _no_dropout(m):
    for s in m.modules():
        if isinstance(s, torch.nn.Dropout):
            s.p = 0.0


def _fill_engine(eng):
    fill_module(eng.network, scale=0.7)
    for name, m in eng.aux.items():
        if name == "Graph":
            fill_module(m)
        elif name == "Grapher":
            fill_module(m, prefix="grapher.")
        elif name == "TGCN":
            fill_module(m, prefix="tgcn.")
        elif name.startswith("Dis_"):
            fill_module(m.dis, prefix=f"dis_{name[4:].lower()}.")
        _no_dropout(m)


def _engine(cfg, dev):
    with contextlib.redirect_stdout(io.StringIO()):
        eng = UDAEngine(cfg, dev)
    _fill_engine(eng)
    return eng


# (engine parameter, oracle group, oracle key): a spread over the trunk, the head, the graph module, a discriminator
GRAD_PROBES = [("network.conv3.weight", "fpn", "conv3.weight"),
               ("network.latlayer2.weight", "fpn", "latlayer2.weight"),
               ("network.semantic_branch.weight", "fpn", "semantic_branch.weight"),
               ("network.gn2.weight", "fpn", "gn2.weight"),
               ("Graph.node_affinity.fc_M.2.weight", "gm", "node_affinity.fc_M.2.weight"),
               ("Graph.intra_domain_graph.linear_q.weight", "gm", "intra_domain_graph.linear_q.weight"),
               ("Dis_P3.dis.cls_logits.weight", "dis_p3", "cls_logits.weight")]


def _engine_param(eng, path):
    head, rest = path.split(".", 1)
    mod = eng.network if head == "network" else eng.aux[head]
    return dict(mod.named_parameters())[rest]


def _check_grads(eng, P, probes, rtol, extra=()):
    errs = {}
    for path, grp, key in list(probes) + list(extra):
        g = _engine_param(eng, path).grad
        ref = P[grp][key].grad
        assert g is not None and ref is not None, path
        errs[path] = float((g.detach().cpu().float() - ref).norm() / ref.norm().clamp_min(1e-12))
    bad = {k: round(v, 4) for k, v in errs.items() if not v < rtol}
    assert not bad, f"relative gradient errors beyond {rtol}: {bad}  (all: { {k: round(v, 4) for k, v in errs.items()} })"


def test_config2_step_losses_and_gradients_match_the_oracle(dev):
    """fp32, no dropout: every entry of the loss dict of one config-2 step equals the CPU oracle's (which runs the
    network per domain, as the reference), and so do parameter gradients after backward (phased step)."""
    cfg = preset(2, bf16=False, cluster_backend="device", cuda_graphs=False)
    eng = _engine(cfg, dev)
    clips, masks = make_batch(cfg, n_clips=2, frames=3)
    fs, ft, _ = split_streams(clips.to(dev))
    total, losses = eng.train_step(fs, masks.to(dev), ft)
    P = OS.build_params(2, "resnet", grapher=True)
    frames = synth.flatten_clips(clips)
    ns = frames.shape[0] // 2
    ref = OS.forward_losses(P, frames[:ns], masks, frames[ns:], num_classes=2, dropout=0.0, cluster=False)
    sum(ref.values()).backward()
    assert set(losses) == set(ref)
    for k in ref:
        torch.testing.assert_close(losses[k].detach().cpu().float(), ref[k].detach().float(), rtol=5e-3, atol=1e-5,
                                   msg=lambda m, k=k: f"{k}: {m}")
    _check_grads(eng, P, GRAD_PROBES, 4e-2, extra=[("network.back_bone.layer3.1.conv2.weight", "fpn", "back_bone.layer3.1.conv2.weight"),
                                                   ("Grapher.fc2.0.weight", "grapher", "fc2.0.weight")])
    # per-domain BatchNorm: two running-stat updates per step, equal to the oracle's two calls
    bn = eng.network.back_bone.bn1
    assert int(bn.num_batches_tracked) == 2
    torch.testing.assert_close(bn.running_mean.cpu(), P["fpn"]["back_bone.bn1.running_mean"], rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(eng.network.back_bone.layer2[0].bn3.running_var.cpu(),
                               P["fpn"]["back_bone.layer2.0.bn3.running_var"], rtol=1e-3, atol=1e-5)


def test_config3_step_matches_the_oracle(dev, fixed_normal):
    """config 3: resnet @ 256x256, 4 classes, SinkhornDistance(0.1, 5, 'mean') between the node sets.  (Classes
    missing in one domain are hallucinated with random draws: `fixed_normal` makes them identical on both sides.)"""
    cfg = preset(3, bf16=False, cluster_backend="device", cuda_graphs=False)
    eng = _engine(cfg, dev)
    eng.aux["Graph"].with_cluster_update = False          # the oracle runs with cluster=False (plain class means)
    xs, masks, xt = make_frame_batch(cfg, 2, 2)
    total, losses = eng.train_step(xs.to(dev), masks.to(dev), xt.to(dev))
    P = OS.build_params(4, "resnet", grapher=False)
    ref = OS.forward_losses(P, xs, masks, xt, num_classes=4, dropout=0.0, cluster=False, sinkhorn_nodes=True,
                            sinkhorn_weight=cfg.sinkhorn_weight)
    sum(ref.values()).backward()
    assert set(losses) == set(ref) and "sinkhorn_loss" in ref
    for k in ref:
        torch.testing.assert_close(losses[k].cpu().float(), ref[k].detach().float(), rtol=5e-3, atol=1e-5,
                                   msg=lambda m, k=k: f"{k}: {m}")
    _check_grads(eng, P, GRAD_PROBES, 3e-2)


def test_config4_step_matches_the_oracle(dev, fixed_normal):
    """config 4: VGG16 @ 256x256, 3 classes, single-frame streams + the temporal clip branch (third network call,
    graph matching with raw logits as score maps -> every target node hallucinated, TGCN) -- losses and gradients
    incl. TGCN's."""
    cfg = preset(4, clip_frames=2, bf16=False, cluster_backend="device", cuda_graphs=False)
    eng = _engine(cfg, dev)
    # the clip branch's graph-matching call hallucinates from the seed banks the first call just updated: both sides
    # must update them the same way (oracle: cluster=False = plain class means)
    eng.aux["Graph"].with_cluster_update = False
    xs, masks, xt = make_frame_batch(cfg, 2, 1)
    clips, tmasks = make_batch(cfg, n_clips=2, frames=2)
    temporal = temporal_input(clips.to(dev), tmasks.to(dev))
    total, losses = eng.train_step(xs.to(dev), masks.to(dev), xt.to(dev), temporal)
    P = OS.build_params(3, "VGG16", grapher=False, tgcn=True, clip_frames=2)
    ref = OS.forward_losses(P, xs, masks, xt, num_classes=3, backbone="VGG16", dropout=0.0, cluster=False,
                            temporal=(synth.flatten_clips(clips), tmasks, (2, 2)))
    sum(ref.values()).backward()
    assert set(losses) == set(ref) and "temporal_graph_loss" in ref
    for k in ref:
        torch.testing.assert_close(losses[k].cpu().float(), ref[k].detach().float(), rtol=1e-2, atol=1e-5,
                                   msg=lambda m, k=k: f"{k}: {m}")
    probes = [p for p in GRAD_PROBES if not p[0].startswith("network.gn2")] + [
        ("network.back_bone.block_3.3.weight", "fpn", "back_bone.block_3.3.weight"),
        ("TGCN.pos_embed", "tgcn", "pos_embed"),
        ("TGCN.grapher.gconv.nn.0.weight", "tgcn", "grapher.gconv.nn.0.weight"),
        ("TGCN.grapher.MLP.0.weight", "tgcn", "grapher.MLP.0.weight")]
    _check_grads(eng, P, probes, 5e-2)
    assert int(eng.network.back_bone.block_1[1].num_batches_tracked) == 3      # source, target, clips


def test_bf16_training_step_against_the_fp32_oracle(dev):
    """The benched numerics (bf16 autocast convolutions, fp32 graph modules) against the fp32 oracle: every loss of a
    full training step.  Stated tolerances: segmentation and adversarial losses 5e-2 relative (+2e-3 absolute); the
    graph-module losses, which sit behind the node sampler + LayerNorm + attention on bf16 features, 0.15 relative
    (measured: node_loss 8.7e-2)."""
    cfg = preset(2, bf16=True, cluster_backend="device", cuda_graphs=False)
    eng = _engine(cfg, dev)
    clips, masks = make_batch(cfg, n_clips=2, frames=3)
    fs, ft, _ = split_streams(clips.to(dev))
    total, losses = eng.train_step(fs, masks.to(dev), ft)
    P = OS.build_params(2, "resnet", grapher=True)
    frames = synth.flatten_clips(clips)
    ns = frames.shape[0] // 2
    ref = OS.forward_losses(P, frames[:ns], masks, frames[ns:], num_classes=2, dropout=0.0, cluster=False)
    for k in ref:
        a, b = float(losses[k]), float(ref[k])
        if k in ("mat_loss_aff", "mat_loss_qu"):
            # tiny (1e-3) matching terms behind a DISCRETE step: the target boxes come from thresholded bf16 logits, a box
            # that moves by one pixel changes which locations are sampled as nodes -- same order of magnitude only
            assert b / 3 <= a <= 3 * b, (k, a, b)
            continue
        graph = k in ("node_loss", "dis_loss")
        torch.testing.assert_close(losses[k].detach().cpu().float(), ref[k].detach().float(), rtol=0.15 if graph else 5e-2,
                                   atol=2e-3, msg=lambda m, k=k: f"{k}: {m}")


def test_bf16_gradients_are_as_accurate_as_stock_autocast(dev):
    """What bf16 costs, and that our kernels cost no more than PyTorch's own: gradients of the segmentation loss through
    (a) the fp32 oracle, (b) the SAME oracle (torch ops only) under torch.autocast(bfloat16), (c) our bf16 path.
    On this network (train-mode BatchNorm / GroupNorm(C,C) mean-subtraction behind bf16 convolutions, noise images) bf16
    rounding is amplified into large gradient errors -- stock autocast: 2 % on conv3, 40-60 % on the lateral layers,
    > 100 % on the stem at 8 frames (profiles/r2_bf16_gradient_snr.md) -- so a fixed tolerance against fp32 is
    meaningless; the criterion is: for every probe our error is at most 1.25 x the stock-autocast error + 0.02, and the
    well-conditioned probes (conv3, gn1) are within 5 % of fp32."""
    from oracle import fpn_ops as FP
    cfg = preset(2, bf16=True, graph_matching=False, vig_grapher=False, cuda_graphs=False)
    clips, masks = make_batch(cfg, n_clips=2, frames=8)
    frames = synth.flatten_clips(clips).to(dev)
    ns, md = frames.shape[0] // 2, masks.to(dev)
    names = ["conv3.weight", "gn1.weight", "semantic_branch.weight", "conv2.weight", "gn2.weight", "smooth3.weight",
             "latlayer2.weight", "toplayer.weight", "back_bone.layer4.2.conv3.weight", "back_bone.conv1.weight"]

    def oracle_grads(autocast):
        P = {k: v.to(dev).detach().requires_grad_(v.requires_grad)
             for k, v in OS.build_params(2, "resnet", grapher=False)["fpn"].items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            logits, _ = FP.fpn_forward(frames[:ns], P, "resnet", training=True)
        FP.seg_loss(logits.float(), md).backward()
        return {n: P[n].grad.float() for n in names}

    eng = _engine(cfg, dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, _ = eng.network(frames[:ns])
    eng.seg_loss(logits, md).backward()
    ours = {n: p.grad.float() for n, p in eng.network.named_parameters() if n in names}
    g32, gac = oracle_grads(False), oracle_grads(True)
    rel = lambda g, r: float((g - r).norm() / r.norm())
    table = {n: (round(rel(gac[n], g32[n]), 4), round(rel(ours[n], g32[n]), 4)) for n in names}
    bad = {n: v for n, v in table.items() if not v[1] <= 1.25 * v[0] + 0.02}
    assert not bad, f"(stock autocast error, our error) vs fp32: {bad}   all: {table}"
    assert table["conv3.weight"][1] < 0.05 and table["gn1.weight"][1] < 0.05, table


def test_training_trajectory_matches_the_oracle(dev):
    """Six optimizer steps on a fixed batch, fp32 eager: the segmentation loss follows the CPU oracle's trajectory
    (Adam on the network, SGD on the rest) step for step, and falls.  Tolerance 6e-2 relative: the first two steps agree
    to 5e-4; after that Adam's sign-like updates amplify the fp32 summation-order differences of 50 train-mode BatchNorm
    layers on 2-frame batches (measured 3e-2 at step 4)."""
    cfg = preset(2, bf16=False, cluster_backend="device", cuda_graphs=False)
    eng = _engine(cfg, dev)
    clips, masks = make_batch(cfg, n_clips=2, frames=2)
    fs, ft, _ = split_streams(clips.to(dev))
    md = masks.to(dev)
    P = OS.build_params(2, "resnet", grapher=True)
    opt = OS.build_optimizers(P)
    frames = synth.flatten_clips(clips)
    ns = frames.shape[0] // 2
    ours, ref = [], []
    for _ in range(6):
        ours.append(float(eng.train_step(fs, md, ft)[1]["seg_loss"]))
        ref.append(float(OS.train_step(P, opt, frames[:ns], masks, frames[ns:], num_classes=2, dropout=0.0,
                                       cluster=False)[1]["seg_loss"]))
    for i, (a, b) in enumerate(zip(ours, ref)):
        assert abs(a - b) <= (0.06 if i > 1 else 1e-2) * abs(b) + 1e-4, (i, ours, ref)
    assert ours[-1] < 0.8 * ours[0], ours


def test_phased_backward_equals_single_graph_backward(dev):
    """The step issues the discriminators' forward+backward before the host-driven GModule (autograd cut at
    the pyramid, gradients joined by hand): losses and every parameter gradient must equal those of the plain
    forward_losses() + one backward()."""
    flats, totals = [], []
    for phased in (False, True):
        cfg = preset(2, bf16=False, cluster_backend="device", cuda_graphs=False, phased_backward=phased, seed=5)
        eng = _engine(cfg, dev)
        clips, masks = make_batch(cfg, n_clips=2, frames=3)
        fs, ft, _ = split_streams(clips.to(dev))
        torch.manual_seed(11)
        total, losses = eng.train_step(fs, masks.to(dev), ft)
        torch.cuda.synchronize()
        flats.append(eng.grads.pack().clone())
        totals.append((total, losses))
    (t0, l0), (t1, l1) = totals
    assert set(l0) == set(l1)
    for k in l0:
        torch.testing.assert_close(l1[k], l0[k], rtol=1e-5, atol=1e-7, msg=lambda m, k=k: f"{k}: {m}")
    torch.testing.assert_close(t1, t0, rtol=1e-5, atol=1e-6)
    rel = (flats[1] - flats[0]).norm() / flats[0].norm()
    assert rel < 1e-4, float(rel)
    assert float(flats[0].abs().sum()) > 0


@pytest.mark.parametrize("bf16", [False, True])
def test_cuda_graph_step_equals_eager_step(dev, bf16):
    """CUDA-graph replay vs eager execution of the same step from the same state: every loss, the whole flat
    gradient, and every buffer (BatchNorm running statistics, num_batches_tracked, seed banks) after the step --
    capture must leave no trace in the model state."""
    res = []
    for graphs in (False, True):
        cfg = preset(2, bf16=bf16, cluster_backend="device", cuda_graphs=graphs, seed=3)
        eng = _engine(cfg, dev)
        clips, masks = make_batch(cfg, n_clips=2, frames=4)
        fs, ft, _ = split_streams(clips.to(dev))
        torch.manual_seed(21)
        total, losses = eng.train_step(fs, masks.to(dev), ft)
        torch.cuda.synchronize()
        eng.aux["Graph"].state_dict()                           # joins the seed stream
        bufs = {n: b.detach().clone() for mod in [eng.network, *eng.aux.values()] for n, b in mod.named_buffers()}
        res.append((total, losses, eng.grads.pack().clone(), bufs))
    (t0, l0, g0, b0), (t1, l1, g1, b1) = res
    tol = 2e-2 if bf16 else 1e-4
    assert set(l0) == set(l1)
    for k in l0:
        torch.testing.assert_close(l1[k], l0[k], rtol=tol, atol=tol * 0.1, msg=lambda m, k=k: f"{k}: {m}")
    rel = (g1 - g0).norm() / g0.norm()
    assert rel < (5e-2 if bf16 else 1e-3), float(rel)
    assert set(b0) == set(b1)
    for n in b0:
        if b0[n].dtype == torch.long:
            assert torch.equal(b0[n], b1[n]), n
        else:
            torch.testing.assert_close(b1[n], b0[n], rtol=tol, atol=tol * 0.1, msg=lambda m, n=n: f"{n}: {m}")


def test_training_makes_progress_and_updates_every_module(dev):
    """The shipped configuration (bf16 + CUDA graphs + phased backward) on a fixed batch: the segmentation loss
    falls steadily -- mean of the last three steps < 0.9 x mean of the first three -- and every module's parameters
    move."""
    cfg = preset(2, bf16=True, cluster_backend="device", cuda_graphs=True)
    eng = _engine(cfg, dev)
    before = {n: next(m.parameters()).detach().clone() for n, m in [("Net", eng.network), *eng.aux.items()]}
    clips, masks = make_batch(cfg, n_clips=2, frames=4)
    fs, ft, _ = split_streams(clips.to(dev))
    seg = []
    for _ in range(12):
        total, losses = eng.train_step(fs, masks.to(dev), ft)
        seg.append(float(losses["seg_loss"]))
        assert all(torch.isfinite(v) for v in losses.values()), losses
    assert sum(seg[-3:]) < 0.9 * sum(seg[:3]), seg
    for n, m in [("Net", eng.network), *eng.aux.items()]:
        assert not torch.equal(next(m.parameters()).detach(), before[n]), n


def test_default_initialisation_also_trains(dev):
    """Same criterion from the modules' own random initialisation with dropout on (what bench.py runs)."""
    cfg = preset(2, bf16=True, cluster_backend="device", cuda_graphs=True)
    with contextlib.redirect_stdout(io.StringIO()):
        eng = UDAEngine(cfg, dev)
    clips, masks = make_batch(cfg, n_clips=2, frames=4)
    fs, ft, _ = split_streams(clips.to(dev))
    seg = [float(eng.train_step(fs, masks.to(dev), ft)[1]["seg_loss"]) for _ in range(20)]
    assert sum(seg[-3:]) < 0.9 * sum(seg[:3]), seg
