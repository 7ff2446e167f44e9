"""One bench step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (numbers printed
under ncu are never bench values).  argv[1] = BASELINE config (2, 3, 4, 5)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from graphecho_b200.engine import UDAEngine, preset
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = bench.WORK[cid]
cfg = preset(cid, cluster_backend="device", cuda_graphs=False, **({"clip_frames": w["frames"]} if cid == 4 else {}))
eng = UDAEngine(cfg, dev)
devin = {k: v.to(dev) for k, v in bench.host_inputs(cfg, w).items()}
def step():
    return eng.train_step(*bench.step_args(devin))[0]
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
