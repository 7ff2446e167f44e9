"""Per-stream kernel counts / summed durations of one config-2 step (torch.profiler chrome trace, CUDA graphs on):
which stream carries how many launches -- the graph module's host-driven side stream vs the graph-replayed trunk.
Diagnostic only (numbers under a profiler are never bench values)."""
import sys, json, collections, gzip
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import profile, ProfilerActivity
from graphecho_b200.engine import EngineConfig, UDAEngine, make_batch, split_streams
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cfg = EngineConfig(hw=112, num_classes=2, bf16=True, cluster_backend="device", cuda_graphs=True)
eng = UDAEngine(cfg, dev)
clips, masks = make_batch(cfg, 8, 32)
clips, masks = clips.to(dev), masks.to(dev)
def step():
    fs, ft, shape = split_streams(clips)
    return eng.train_step(fs, masks, ft)[0]
for _ in range(6): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
path = Path("gpurun_out/trace_streams.json")
prof.export_chrome_trace(str(path))
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
t0 = min(e["ts"] for e in ks)
by = collections.defaultdict(list)
for e in ks:
    by[e["args"].get("stream")].append(e)
for s, lst in sorted(by.items(), key=lambda kv: -len(kv[1])):
    dur = sum(e["dur"] for e in lst)
    first, last = min(e["ts"] for e in lst) - t0, max(e["ts"] + e["dur"] for e in lst) - t0
    print(f"stream {s}: {len(lst)} launches, {dur/1e3:.2f} ms busy, active {first/1e3:.2f}..{last/1e3:.2f} ms")
    names = collections.Counter(e["name"][:70] for e in lst)
    durs = collections.defaultdict(float)
    for e in lst:
        durs[e["name"][:70]] += e["dur"]
    for n, c in names.most_common(14):
        print(f"      {c:4d} x {durs[n]/1e3:7.3f} ms  {n}")
# CPU side: how long the host spends inside the graph module forward and in each backward() call
cpu = [e for e in ev if e.get("cat") in ("cpu_op", "user_annotation", "python_function")]
print("cuda launches issued by the host (cudaLaunchKernel etc.):",
      sum(1 for e in ev if e.get("cat") == "cuda_runtime" and "Launch" in e.get("name", "")))
path.unlink()
