"""CPU: oracle/step.py (the restatement, the step-level parity checker of the GPU tests) is pinned to a step
composed of the UNMODIFIED reference modules in the reference trainer's order (oracle/ref_step.py over
oracle/_ref, staged from /root/reference by oracle/build_ref.py).  Skipped where the reference is not staged."""
import warnings

import pytest
import torch

from graphecho_b200 import synth
from graphecho_b200.engine import make_batch, make_frame_batch, preset
from oracle import ref_step as RS
from oracle import step as OS

pytestmark = pytest.mark.skipif(RS.reference_root() is None, reason="reference not staged (oracle/_ref) nor mounted")


def _compare(ref, ours, rtol=2e-4, atol=2e-6):
    assert set(ref) == set(ours), (sorted(ref), sorted(ours))
    for k in ref:
        torch.testing.assert_close(ours[k].detach(), ref[k].detach(), rtol=rtol, atol=atol, msg=lambda m, k=k: f"{k}: {m}")


def test_config2_step_matches_reference_modules():
    """FPN(resnet) + Grapher(p2) + GModule (with sklearn seed update) + 4 discriminators, two steps: losses, and the
    state every step leaves behind (BatchNorm running statistics of a layer, the seed banks)."""
    warnings.filterwarnings("ignore")
    cfg = preset(2)
    clips, masks = make_batch(cfg, n_clips=2, frames=2)
    frames = synth.flatten_clips(clips)
    ns = frames.shape[0] // 2
    torch.manual_seed(0)
    M = RS.build_modules(2, "resnet", grapher=True, dropout=0.0)
    P = OS.build_params(2, "resnet", grapher=True)
    opt_r, opt_o = RS.build_optimizers(M), OS.build_optimizers(P)
    for step in range(2):
        torch.manual_seed(10 + step)
        tr, lr_ = RS.train_step(M, opt_r, frames[:ns], masks, frames[ns:])
        torch.manual_seed(10 + step)
        to, lo = OS.train_step(P, opt_o, frames[:ns], masks, frames[ns:], num_classes=2, dropout=0.0, cluster=True)
        _compare(lr_, lo, rtol=2e-3 if step else 2e-4, atol=1e-5)
    sd = M["net"].state_dict()
    for name in ("back_bone.bn1.running_mean", "back_bone.layer2.0.bn3.running_var"):
        torch.testing.assert_close(P["fpn"][name].float(), sd[name].float(), rtol=1e-3, atol=1e-5)
    assert int(sd["back_bone.bn1.num_batches_tracked"]) == 4            # two network calls per step
    torch.testing.assert_close(P["gm"]["sr_seed"], M["gm"].sr_seed, rtol=1e-3, atol=1e-4)


def test_config3_and_4_losses_match_reference_modules():
    warnings.filterwarnings("ignore")
    # config 3: 256x256, nc=4, SinkhornDistance between the node sets
    cfg = preset(3)
    xs, masks, xt = make_frame_batch(cfg, 2, 2)
    M = RS.build_modules(4, "resnet", grapher=False, dropout=0.0)
    P = OS.build_params(4, "resnet", grapher=False)
    torch.manual_seed(5)        # classes missing in one domain are hallucinated with torch.normal: same stream both sides
    ref = RS.forward_losses(M, xs, masks, xt, sinkhorn_nodes=True, sinkhorn_weight=cfg.sinkhorn_weight)
    torch.manual_seed(5)
    ours = OS.forward_losses(P, xs, masks, xt, num_classes=4, dropout=0.0, cluster=True, sinkhorn_nodes=True,
                             sinkhorn_weight=cfg.sinkhorn_weight)
    _compare(ref, ours)
    # config 4: VGG16, nc=3, temporal clips + TGCN (2 clips x 2 frames)
    cfg = preset(4, clip_frames=2)
    xs, masks, xt = make_frame_batch(cfg, 2, 1)
    clips, tmasks = make_batch(cfg, n_clips=2, frames=2)
    temporal = (synth.flatten_clips(clips), tmasks, (2, 2))
    M = RS.build_modules(3, "VGG16", grapher=False, tgcn=True, clip_frames=2, dropout=0.0)
    P = OS.build_params(3, "VGG16", grapher=False, tgcn=True, clip_frames=2)
    torch.manual_seed(6)
    ref = RS.forward_losses(M, xs, masks, xt, temporal=temporal)
    torch.manual_seed(6)
    ours = OS.forward_losses(P, xs, masks, xt, num_classes=3, backbone="VGG16", dropout=0.0, cluster=True, temporal=temporal)
    _compare(ref, ours)
