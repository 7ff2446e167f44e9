"""Debug: config-4 gradient probes vs the oracle, decomposed (main pass only / with temporal)."""
import sys, contextlib, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from graphecho_b200 import synth
from graphecho_b200.engine import preset, make_batch, make_frame_batch, temporal_input, UDAEngine
from oracle import step as OS
import test_engine_gpu as T
import conftest

dev = torch.device("cuda:0")
gen = conftest.fixed_normal.__wrapped__() if hasattr(conftest.fixed_normal, "__wrapped__") else None
if gen is not None:
    next(gen)
for with_temporal in (False, True):
    for phased in (True, False):
        cfg = preset(4, clip_frames=2, bf16=False, cluster_backend="device", cuda_graphs=False, phased_backward=phased)
        eng = T._engine(cfg, dev)
        eng.aux["Graph"].with_cluster_update = False
        xs, masks, xt = make_frame_batch(cfg, 2, 1)
        clips, tmasks = make_batch(cfg, n_clips=2, frames=2)
        temporal = temporal_input(clips.to(dev), tmasks.to(dev)) if with_temporal else None
        total, losses = eng.train_step(xs.to(dev), masks.to(dev), xt.to(dev), temporal)
        P = OS.build_params(3, "VGG16", grapher=False, tgcn=True, clip_frames=2)
        ref = OS.forward_losses(P, xs, masks, xt, num_classes=3, backbone="VGG16", dropout=0.0, cluster=False,
                                temporal=(synth.flatten_clips(clips), tmasks, (2, 2)) if with_temporal else None)
        sum(ref.values()).backward()
        print(f"--- temporal={with_temporal} phased={phased}")
        for k in ref:
            print(f"  {k:22s} ours {float(losses[k]):.6f} ref {float(ref[k]):.6f}")
        probes = T.GRAD_PROBES + [("network.back_bone.block_3.3.weight", "fpn", "back_bone.block_3.3.weight"),
                                  ("network.latlayer1.weight", "fpn", "latlayer1.weight"), ("network.latlayer3.weight", "fpn", "latlayer3.weight"),
                                  ("network.toplayer.weight", "fpn", "toplayer.weight"), ("network.smooth2.weight", "fpn", "smooth2.weight")]
        if with_temporal:
            probes += [("TGCN.pos_embed", "tgcn", "pos_embed"), ("TGCN.grapher.gconv.nn.0.weight", "tgcn", "grapher.gconv.nn.0.weight"),
                       ("TGCN.grapher.MLP.0.weight", "tgcn", "grapher.MLP.0.weight"), ("TGCN.graph_attention.linear_k.weight", "tgcn", "graph_attention.linear_k.weight")]
        for path, grp, key in probes:
            g = T._engine_param(eng, path).grad
            r = P[grp][key].grad
            if g is None or r is None:
                print(f"  {path}: grad None ours={g is None} ref={r is None}")
                continue
            err = (g.detach().cpu().float() - r).norm() / r.norm().clamp_min(1e-12)
            print(f"  {path:45s} rel err {float(err):.3e}   |ref| {float(r.norm()):.3e}")
