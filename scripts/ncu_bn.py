import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import functional as GF, _cabi
dev = torch.device("cuda:0")
cl = torch.channels_last
path = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_cabi.lib().ge_bn_set_path(path)
shapes = [(256, 256, 28, True), (256, 64, 56, False), (256, 256, 14, False), (256, 512, 4, False)]
items = []
for N, C, H, res in shapes:
    x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_()
    r = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_() if res else None
    bn = torch.nn.BatchNorm2d(C).to(dev)
    g = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl)
    items.append((x, r, bn, g))
def run():
    for x, r, bn, g in items:
        with GF.domain_split(x.shape[0] // 2):
            o = GF.bn_act(x, bn, residual=r, relu=True)
        torch.autograd.grad(o, [x] + ([r] if r is not None else []), g)
run(); torch.cuda.synchronize()
torch.cuda.profiler.start()
run(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
