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
    return eng.train_step(fs, masks, ft)[0]
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
from torch.autograd import DeviceType
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for evt in prof.events():
    if evt.device_type == DeviceType.CUDA:
        agg[evt.name][0] += 1
        agg[evt.name][1] += evt.device_time_total if hasattr(evt, "device_time_total") else evt.cuda_time_total
tot = sum(v[1] for v in agg.values())
print("GPU kernel ms per step", tot / 2e3, "kernels per step", sum(v[0] for v in agg.values()) // 2)
for name, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{v[1]/2e3:8.3f} ms {100*v[1]/tot:5.1f}% {v[0]//2:5d}  {name[:120]}")
import time
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
