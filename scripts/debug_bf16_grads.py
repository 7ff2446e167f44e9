"""Is the bf16 gradient error of the head parameters inherent to bf16 autocast (stock PyTorch shows it too) or ours?"""
import sys, contextlib, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from graphecho_b200 import synth
from graphecho_b200.engine import preset, make_batch, UDAEngine, split_streams
from graphecho_b200 import functional as GF
from oracle import step as OS, fpn_ops as FP
from oracle.detfill import fill_module
dev = torch.device("cuda:0")
cfg = preset(2, bf16=True, graph_matching=False, vig_grapher=False, cuda_graphs=False)
FR = int(sys.argv[1]) if len(sys.argv) > 1 else 3
clips, masks = make_batch(cfg, n_clips=2, frames=FR)
frames = synth.flatten_clips(clips).to(dev)
ns = frames.shape[0] // 2
md = masks.to(dev)
names = ["conv3.weight", "semantic_branch.weight", "conv2.weight", "gn1.weight", "gn2.weight", "gn2.bias", "smooth3.weight", "latlayer2.weight",
         "latlayer3.weight", "toplayer.weight", "back_bone.layer4.2.conv3.weight", "back_bone.layer1.0.conv1.weight", "back_bone.conv1.weight"]

def oracle_grads(autocast):
    P = {k: v.to(dev).detach().requires_grad_(v.requires_grad) for k, v in OS.build_params(2, "resnet", grapher=False)["fpn"].items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits, _ = FP.fpn_forward(frames[:ns], P, "resnet", training=True)
    loss = FP.seg_loss(logits.float(), md)
    loss.backward()
    return float(loss), {n: P[n].grad.float().clone() for n in names}

def our_grads(bf16):
    c = preset(2, bf16=bf16, graph_matching=False, vig_grapher=False, cuda_graphs=False)
    with contextlib.redirect_stdout(io.StringIO()):
        eng = UDAEngine(c, dev)
    fill_module(eng.network, scale=0.7)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
        logits, _ = eng.network(frames[:ns])
    loss = eng.seg_loss(logits, md)
    loss.backward()
    sd = dict(eng.network.named_parameters())
    return float(loss), {n: sd[n].grad.float().clone() for n in names}

l32, g32 = oracle_grads(False)
lac, gac = oracle_grads(True)
lo32, go32 = our_grads(False)
lobf, gobf = our_grads(True)
print(f"frames per domain {FR}; loss: oracle fp32 {l32:.5f}  stock autocast {lac:.5f}  ours fp32 {lo32:.5f}  ours bf16 {lobf:.5f}")
print(f"{'param':40s} {'stock-autocast':>15s} {'ours-fp32':>12s} {'ours-bf16':>12s}")
for n in names:
    r = g32[n]
    e = lambda g: float((g - r).norm() / r.norm())
    print(f"{n:40s} {e(gac[n]):15.4f} {e(go32[n]):12.5f} {e(gobf[n]):12.4f}")
