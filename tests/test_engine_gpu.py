"""GPU: the whole training step (graphecho_b200.engine.UDAEngine) — step-level parity with the CPU
oracle step (oracle/step.py: FPN + Grapher + GModule + 4 discriminators, same name-keyed weights),
CUDA-graph replay vs eager execution, the TGCN temporal configuration, and optimizer progress."""
import contextlib
import io

import pytest
import torch

from graphecho_b200 import synth
from graphecho_b200.engine import EngineConfig, UDAEngine, make_batch, split_streams
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
        elif name.startswith("Dis_"):
            fill_module(m.dis, prefix=f"dis_{name[4:].lower()}.")
        _no_dropout(m)


def test_step_losses_match_the_oracle_step(dev):
    """fp32, no dropout, sklearn-free: every entry of the loss dict of one step equals the CPU oracle's."""
    cfg = EngineConfig(hw=112, num_classes=2, bf16=False, cluster_backend="device", cuda_graphs=False)
    with contextlib.redirect_stdout(io.StringIO()):
        eng = UDAEngine(cfg, dev)
    _fill_engine(eng)
    clips, masks = make_batch(cfg, n_clips=2, frames=3)
    fs, ft, shape = split_streams(clips.to(dev))
    losses = eng.forward_losses(fs, masks.to(dev), ft, shape)
    P = OS.build_params(2, "resnet", grapher=True)
    frames = synth.flatten_clips(clips)
    ns = frames.shape[0] // 2
    ref = OS.forward_losses(P, frames[:ns], masks, frames[ns:], num_classes=2, dropout=0.0, cluster=False)
    assert set(losses) == set(ref)
    for k in ref:
        torch.testing.assert_close(losses[k].detach().cpu().float(), ref[k].detach().float(), rtol=5e-3, atol=1e-5,
                                   msg=lambda m, k=k: f"{k}: {m}")


def test_phased_backward_equals_single_graph_backward(dev):
    """The step issues the discriminators' forward+backward before the host-driven GModule (autograd cut at
    the pyramid, gradients joined by hand): losses and every parameter gradient must equal those of the plain
    forward_losses() + one backward()."""
    flats, totals = [], []
    for phased in (False, True):
        cfg = EngineConfig(hw=112, num_classes=2, bf16=False, cluster_backend="device", cuda_graphs=False,
                           phased_backward=phased, seed=5)
        with contextlib.redirect_stdout(io.StringIO()):
            eng = UDAEngine(cfg, dev)
        _fill_engine(eng)
        clips, masks = make_batch(cfg, n_clips=2, frames=3)
        fs, ft, shape = split_streams(clips.to(dev))
        torch.manual_seed(11)
        total, losses = eng.train_step(fs, masks.to(dev), ft, shape)
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
    res = []
    for graphs in (False, True):
        cfg = EngineConfig(hw=112, num_classes=2, bf16=bf16, cluster_backend="device", cuda_graphs=graphs, seed=3)
        with contextlib.redirect_stdout(io.StringIO()):
            eng = UDAEngine(cfg, dev)
        for m in eng.aux.values():
            _no_dropout(m)
        clips, masks = make_batch(cfg, n_clips=2, frames=4)
        fs, ft, shape = split_streams(clips.to(dev))
        out = [eng.train_step(fs, masks.to(dev), ft, shape) for _ in range(3)]
        res.append(out)
    tol = 5e-2 if bf16 else 2e-3
    # step 0 is bit-for-bit the same computation; later steps inherit the (tiny) differences of the updates
    for (t0, l0), (t1, l1) in zip(*res):
        assert torch.isfinite(t0) and torch.isfinite(t1)
        assert set(l0) == set(l1)
        torch.testing.assert_close(t1, t0, rtol=tol, atol=tol)


def test_training_makes_progress_and_updates_every_module(dev):
    cfg = EngineConfig(hw=112, num_classes=2, bf16=True, cluster_backend="device", cuda_graphs=True)
    with contextlib.redirect_stdout(io.StringIO()):
        eng = UDAEngine(cfg, dev)
    before = {n: next(m.parameters()).detach().clone() for n, m in [("Net", eng.network), *eng.aux.items()]}
    clips, masks = make_batch(cfg, n_clips=2, frames=4)
    fs, ft, shape = split_streams(clips.to(dev))
    seg = []
    for _ in range(12):
        total, losses = eng.train_step(fs, masks.to(dev), ft, shape)
        seg.append(float(losses["seg_loss"]))
        assert all(torch.isfinite(v) for v in losses.values()), losses
    # the segmentation loss gets below its starting value (single steps are noisy: 8 frames, float atomics, bf16)
    assert min(seg[1:]) < seg[0], seg
    for n, m in [("Net", eng.network), *eng.aux.items()]:
        assert not torch.equal(next(m.parameters()).detach(), before[n]), n


def test_temporal_configuration_vgg16_tgcn(dev):
    """Config-4 shape family: VGG16 backbone, 256x256 clips, 3 classes, TGCN temporal module on."""
    cfg = EngineConfig(backbone="VGG16", hw=256, num_classes=3, bf16=True, vig_grapher=False, temporal_graph=True,
                       clip_frames=4, cluster_backend="device", cuda_graphs=False)
    with contextlib.redirect_stdout(io.StringIO()):
        eng = UDAEngine(cfg, dev)
    clips, masks = make_batch(cfg, n_clips=2, frames=4)
    fs, ft, shape = split_streams(clips.to(dev))
    total, losses = eng.train_step(fs, masks.to(dev), ft, shape)
    assert torch.isfinite(total)
    assert "temporal_graph_loss" in losses and "mat_loss_aff" in losses and "loss_adv_p5" in losses
    g = eng.aux["TGCN"].pos_embed.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
