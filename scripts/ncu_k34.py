"""Affinity (K3) and Sinkhorn (K4) kernels, one problem and 512 problems of 252 x 252 (x 512 hidden), forward and
backward, between cudaProfilerStart/Stop for `ncu --set full --import-source on --profile-from-start off`."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import functional as GF
dev = torch.device("cuda:0")
torch.manual_seed(0)
w2, b2 = torch.randn(512, device=dev), torch.randn(1, device=dev)


def run(batch, n=252):
    A = torch.randn(batch, n, 512, device=dev).requires_grad_()
    B = torch.randn(batch, n, 512, device=dev).requires_grad_()
    M = GF.affinity_pairwise(A, B, w2, b2)
    Md = M.detach().requires_grad_()
    P = GF.sinkhorn_rpm_exp(Md, 20, True)
    P.backward(torch.randn_like(P))
    M.backward(Md.grad)


for b in (1, 512):
    run(b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for b in (1, 512):
    run(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
