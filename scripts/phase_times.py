"""Wall-clock / CUDA-event phase timing of the phased step without a profiler attached (diagnostic)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import engine as E
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
# monkeypatch-free: time the GModule call on the CPU and with events on the side stream
gm = eng._gmodule
orig = gm.forward_joint
marks = {}
def timed_fj(*a, **k):
    s = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
    t0 = time.perf_counter()
    out = orig(*a, **k)
    marks["gm_fwd_cpu_ms"] = (time.perf_counter() - t0) * 1e3
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s)
    marks["ev"] = (e0, e1)
    return out
gm.forward_joint = timed_fj
orig_bw = torch.autograd.backward
def timed_bw(*a, **k):
    t0 = time.perf_counter()
    r = orig_bw(*a, **k)
    marks.setdefault("bw_cpu_ms", []).append((time.perf_counter() - t0) * 1e3)
    ev = torch.cuda.Event(enable_timing=True); ev.record(torch.cuda.current_stream())
    marks.setdefault("bw_ev", []).append(ev)
    return r
torch.autograd.backward = timed_bw
for it in range(8):
    marks.clear()
    torch.cuda.synchronize()
    ev_a = torch.cuda.Event(enable_timing=True); ev_b = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); ev_a.record()
    step()
    t_cpu = (time.perf_counter() - t0) * 1e3
    ev_b.record(); torch.cuda.synchronize()
    e0, e1 = marks["ev"]
    bev = marks["bw_ev"]
    print("backward() calls end at (ms into step, on their stream):", [round(ev_a.elapsed_time(e), 2) for e in bev])
    print(f"step gpu {ev_a.elapsed_time(ev_b):.2f} ms, cpu issue {t_cpu:.2f} ms | GModule fwd: cpu {marks['gm_fwd_cpu_ms']:.2f} ms, side-stream {e0.elapsed_time(e1):.2f} ms"
          f" (starts {ev_a.elapsed_time(e0):.2f} ms into the step) | backward() cpu ms: {[round(x,2) for x in marks['bw_cpu_ms']]}")
