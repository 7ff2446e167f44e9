"""torch.profiler breakdown of one bench step (diagnostic; numbers under a profiler are never bench values)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import profile, ProfilerActivity
from graphecho_b200.engine import EngineConfig, UDAEngine, make_batch, split_streams
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cfg = EngineConfig(hw=112, num_classes=2, bf16=True, cluster_backend="device")
eng = UDAEngine(cfg, dev)
clips, masks = make_batch(cfg, 8, 32)
clips, masks = clips.to(dev), masks.to(dev)
def step():
    fs, ft, shape = split_streams(clips)
    return eng.train_step(fs, masks, ft, shape)[0]
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print("total self device ms per step", tot / 2e3)
for e in rows[:45]:
    print(f"{e.self_device_time_total/2e3:8.3f} ms {100*e.self_device_time_total/tot:5.1f}% {e.count//2:5d}  {e.key[:110]}")
import time
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
