"""One call of every custom kernel family at its config-2 shape between cudaProfilerStart/Stop, for
`ncu --set full --profile-from-start off` (DRAM traffic / pipe utilisation per launch)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import functional as GF
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, C, H = 256, 256, 28
cl = torch.channels_last
x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl)
res = torch.randn_like(x)
bn = torch.nn.BatchNorm2d(C).to(dev)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
xn = torch.randn(N, H * H, C, device=dev).bfloat16()
A, B = torch.randn(1, 252, 512, device=dev), torch.randn(1, 252, 512, device=dev)
w2, b2 = torch.randn(512, device=dev), torch.randn(1, device=dev)
M = torch.randn(1, 252, 252, device=dev)
s = [torch.randn(N, 128, H, H, device=dev).bfloat16().contiguous(memory_format=cl) for _ in range(4)]
W3, b3 = torch.randn(2, 128, 1, 1, device=dev), torch.randn(2, device=dev)


def run():
    xr = x.clone().requires_grad_()
    o = GF.bn_act(xr, bn, residual=res, relu=True)
    o.backward(torch.ones_like(o))
    xg = x.clone().requires_grad_()
    o = GF.gn_relu(xg, gamma, beta, 32)
    o.backward(torch.ones_like(o))
    e = GF.knn_graph_nmajor(xn, None, 9, 1)[0]
    xr2 = xn.clone().requires_grad_()
    f = GF.mr_gather_nmajor(xr2, e)
    f.backward(torch.ones_like(f))
    Ar, Br = A.clone().requires_grad_(), B.clone().requires_grad_()
    Mx = GF.affinity_pairwise(Ar, Br, w2, b2)
    Mx.sum().backward()
    Mr = M.clone().requires_grad_()
    P = GF.sinkhorn_rpm_exp(Mr, 20, True)
    P.sum().backward()
    GF.seg_tail(*s, W3, b3, 4)


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
