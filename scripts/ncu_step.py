"""One bench step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (numbers printed
under ncu are never bench values)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200.engine import EngineConfig, UDAEngine, make_batch, split_streams
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cfg = EngineConfig(hw=112, num_classes=2, bf16=True, cluster_backend="device")
eng = UDAEngine(cfg, dev)
clips, masks = make_batch(cfg, 8, 32)
clips, masks = clips.to(dev), masks.to(dev)
def step():
    fs, ft, shape = split_streams(clips)
    return eng.train_step(fs, masks, ft)[0]
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
