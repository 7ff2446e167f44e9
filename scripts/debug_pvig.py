import sys, contextlib, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from graphecho_b200.models import vig
from graphecho_b200 import functional as GF
from oracle import vig_ops as V
from oracle.detfill import fill_module
dev = torch.device("cuda:0")
torch.manual_seed(0)
for C, N, d, r in ((240, 196, 2, 1), (240, 196, 3, 1), (384, 49, 3, 1), (96, 784, 1, 2), (48, 196, 2, 1)):
    H = int(N ** 0.5)
    with contextlib.redirect_stdout(io.StringIO()):
        gr = fill_module(vig.Grapher(C, 9, d, "mr", "gelu", "batch", True, False, 0.2, r, N, 0.0, True), prefix="g.").to(dev).eval()
    P = {k: v.detach().cpu() for k, v in gr.state_dict().items()}
    x = torch.randn(2, C, H, H)
    with torch.no_grad():
        out = gr(x.to(dev))
        ref, e_ref = V.grapher(x, P, "", k=9, dilation=d, r=r, norm="batch", act="gelu", training=False,
                               relative_pos=P["relative_pos"], return_edges=True)
        t = V.conv_bn(x, P, "fc1.", False)
        tt = t.reshape(2, C, -1, 1)
        y = torch.nn.functional.avg_pool2d(t, r, r).reshape(2, C, -1, 1) if r > 1 else None
        e = GF.knn_graph(tt.to(dev), None if y is None else y.to(dev), 9, d, P["relative_pos"].to(dev)).cpu()
        mism = (e[0] != e_ref[0]).float().mean()
        feat = GF.mr_gather(tt.to(dev), e_ref.to(dev), None if y is None else y.to(dev)).cpu()
        fref = V.max_relative(tt, e_ref, y)
    print(f"C{C} N{N} d{d} r{r}: out rel err {float((out.cpu()-ref).norm()/ref.norm()):.3e}  knn mismatch {float(mism):.4f}  "
          f"gather max err {float((feat-fref).abs().max()):.3e}  edge shape {tuple(e.shape)} ref {tuple(e_ref.shape)}")
    if mism > 0.01:
        print("   ours", e[0,0,0].tolist(), "\n   ref ", e_ref[0,0,0].tolist())
        dist = V.knn_distances(tt, y, P["relative_pos"])
        order = dist[0,0].argsort()[:9*d]
        print("   sorted", order.tolist())
