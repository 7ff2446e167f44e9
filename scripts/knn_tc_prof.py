"""One tcgen05 k-NN call at the bench shape, for ncu."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import functional as GF
dev = torch.device("cuda:0")
x = torch.randn(256, 256, 784, 1, device=dev)
for _ in range(2):
    GF.knn_graph(x, None, 9, 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
GF.knn_graph(x, None, 9, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
