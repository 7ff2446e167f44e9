"""Print relative-norm errors of the product modules against the golden fixtures (diagnostic)."""
import contextlib, io, sys
from pathlib import Path
import torch, torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from graphecho_b200 import synth
from graphecho_b200.models import fpnseg, TGCN as tgcn_mod
from graphecho_b200.utils.sinkhorn_distance import SinkhornDistance
from graphecho_b200.utils.losses import DiceLoss
from oracle.detfill import fill_module
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
G = Path(__file__).resolve().parent.parent / "tests" / "golden"
def rel(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return f"rel={((a-b).norm()/b.norm()).item():.2e} maxabs={(a-b).abs().max().item():.2e} refmax={b.abs().max().item():.2e}"
g = torch.load(G / "fpn.pt", weights_only=False)
for bb, nc, hw in (("resnet", 1, 112), ("VGG16", 3, 64)):
    x = g[f"{bb}_x"].to(dev)
    for mode in ("eval", "train"):
        rec = g[f"{bb}_{mode}"]
        net = fill_module(fpnseg.FPN([2, 4, 23, 3], nc, 1, back_bone=bb), scale=0.7).to(dev)
        net.train(mode == "train")
        xr = x.clone().requires_grad_(mode == "train")
        logits, feats = net(xr)
        print(bb, mode, "logits", rel(logits, rec["logits"]), "p5", rel(feats[3], rec["p5"]), "p4", rel(feats[2], rec["p4"]))
        if mode == "train":
            B = x.shape[0]
            mask = (synth.disc_masks(B, nc, hw) if nc > 1 else synth.disc_masks(B, 2, hw)[:, 1:2]).to(dev)
            loss = DiceLoss()(logits, mask) + F.binary_cross_entropy_with_logits(logits, mask)
            loss.backward()
            print("   loss", loss.item(), rec["loss"].item())
            print("   dconv3", rel(net.conv3.weight.grad, rec["dconv3"]))
            print("   dgn1", rel(net.gn1.weight.grad, rec["dgn1"]))
            print("   dsem", rel(net.semantic_branch.weight.grad[:4], rec["dsem"]))
            print("   dtop", rel(net.toplayer.weight.grad[:4, :64], rec["dtop"]))
            print("   dx", rel(xr.grad, rec["dx"]))
t = torch.load(G / "tgcn.pt", weights_only=False)
for transport in ("node_discriminate", "sinkhorn_distance"):
    rec = t[transport]
    with contextlib.redirect_stdout(io.StringIO()):
        m = fill_module(tgcn_mod.TGCN(256, 256, (3, 8, 8), 10, 10, None, transport)).to(dev).train()
    for s in m.modules():
        if isinstance(s, torch.nn.Dropout): s.p = 0.0
    feats = [f.to(dev).requires_grad_() for f in synth.clip_pyramid(2, 3, 256, seed=31)]
    idx = (torch.zeros(1, dtype=torch.long, device=dev),) * 2
    losses = m(feats, (rec["src"].to(dev), rec["tgt"].to(dev)), SinkhornDistance(0.1, 5, "mean"), torch.nn.CrossEntropyLoss(), idx, r=[8, 4, 2, 1])
    sum(losses.values()).backward()
    print(transport, {k: (v.item(), rec["losses"][k].item()) for k, v in losses.items()})
    print("   df3", rel(feats[3].grad, rec["df3"]), "dpos", rel(m.pos_embed.grad[:, :, :8], rec["dpos"]))
