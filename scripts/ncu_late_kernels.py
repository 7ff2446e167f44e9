"""The kernels added at the end of round 2 at their config-2 shapes, between cudaProfilerStart/Stop for
`ncu --set full --profile-from-start off`: stem convolution (forward, weight gradient), fused bias add / bias gradient
on the BatchNorm kernels, node-sampler labels + gather + scatter, matching loss."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import functional as GF
from graphecho_b200.models.graph_matching import GModule, PrototypeComputation
dev = torch.device("cuda:0")
torch.manual_seed(0)
cl = torch.channels_last
conv1 = torch.nn.Conv2d(1, 64, 7, 2, 3, bias=False).to(dev)
x = torch.randn(256, 1, 112, 112, device=dev)
lat = torch.nn.Conv2d(256, 256, 1).to(dev)
c2 = torch.randn(256, 256, 28, 28, device=dev).bfloat16().contiguous(memory_format=cl)
hw = [(28, 28), (14, 14), (7, 7), (4, 4)]
feats = [torch.randn(256, 256, h, w, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_() for h, w in hw]
masks = torch.zeros(128, 2, 112, 112, device=dev)
masks[:, 1, 30:80, 25:90] = 1
masks[:, 0] = 1 - masks[:, 1]
gm = GModule(256, 2, dev).to(dev)
gen = PrototypeComputation(2)
P = GF.sinkhorn_rpm_exp(torch.randn(252, 250, device=dev), 20, True).detach().requires_grad_()
l1, l2 = torch.randint(0, 2, (252,), device=dev).float(), torch.randint(0, 2, (250,), device=dev).float()


def run():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = GF.stem_conv(x, conv1)
        y.backward(torch.ones_like(y))
        z = GF.conv_bias(c2, lat)
        z.backward(torch.ones_like(z))
    boxes = gm.find_bbox(masks)
    labels, counts = gen.plan(None, boxes, hw, gm.fpn_strides)
    cnt = counts.tolist()
    (n, l, w), (n2, l2_, w2) = gen.gather_pair((feats, labels, cnt, 0), (feats, labels, cnt, 128))
    (n.sum() + n2.sum()).backward()
    GF.matching_loss_o2o(P, l1, l2).backward()


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
