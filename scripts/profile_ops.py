"""torch.profiler attribution of one bench step: CUDA time per aten op with input shapes + chrome trace
(diagnostic; numbers under a profiler are never bench values)."""
import sys, json, collections
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import profile, ProfilerActivity
from graphecho_b200.engine import EngineConfig, UDAEngine, make_batch, split_streams
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
graphs = "--graphs" in sys.argv
cfg = EngineConfig(hw=112, num_classes=2, bf16=True, cluster_backend="device", cuda_graphs=graphs)
eng = UDAEngine(cfg, dev)
clips, masks = make_batch(cfg, 8, 32)
clips, masks = clips.to(dev), masks.to(dev)
def step():
    fs, ft, shape = split_streams(clips)
    return eng.train_step(fs, masks, ft)[0]
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
out = Path("gpurun_out"); out.mkdir(exist_ok=True)
tag = "graphs" if graphs else "eager"
prof.export_chrome_trace(str(out / f"trace_{tag}.json"))
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=60, max_shapes_column_width=70))
