#!/usr/bin/env python
"""Benchmark of GraphEcho's data-parallel hot path (BASELINE.json metric: frames/s on 112x112,
32-frame clips; config 2: FPN(resnet) + ViG Grapher + graph matching + 4 discriminators, bf16,
8 clips per GPU).  One "step" = one full training step (forward + backward + optimizer) over one
batch of synthetic clips.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # ours (one rank per GPU under torchrun)
  python bench.py --impl reference ...                          # the CPU oracle port of the same step

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "frames/s (112x112, 32-frame clips), full training step"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=8, help="clips per GPU (half source, half target)")
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--hw", type=int, default=112)
    ap.add_argument("--num-classes", type=int, default=2)
    ap.add_argument("--sync-bn", action="store_true",
                    help="SyncBatchNorm on the network when N>1 (the reference's intent; eager only, no CUDA graphs)")
    ap.add_argument("--no-graphs", action="store_true", help="run the static segments eagerly instead of as CUDA graphs")
    ap.add_argument("--fp32", action="store_true", help="fp32 convolutions instead of bf16 autocast")
    ap.add_argument("--cpu-sample-frames", type=int, default=16, help="frames per clip in the CPU baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"config2: {a.clips} clips x {a.frames} frames {a.hw}x{a.hw} per GPU (half source, half target), "
            f"FPN(resnet,nc={a.num_classes}) + ViG Grapher(p2,k=9) + GModule + 4 Discriminators, fwd+bwd+optim")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, enabled=True):
        self.index, self.proc, self.lines, self.enabled = index, None, [], enabled

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_step_runner(a, frames_per_clip):
    """The oracle port of the same step on the host cores (all threads)."""
    from oracle import step as OS
    from graphecho_b200 import synth
    from graphecho_b200.engine import EngineConfig, make_batch
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = EngineConfig(hw=a.hw, num_classes=a.num_classes)
    clips, masks = make_batch(cfg, n_clips=2, frames=frames_per_clip)
    frames = synth.flatten_clips(clips)
    ns = frames.shape[0] // 2
    P = OS.build_params(a.num_classes, "resnet", grapher=True)
    opt = OS.build_optimizers(P)

    def step():
        total, _ = OS.train_step(P, opt, frames[:ns], masks, frames[ns:], num_classes=a.num_classes)
        return float(total)

    return step, frames.shape[0]


def time_cpu(a, frames_per_clip, steps, warmup):
    step, nframes = cpu_step_runner(a, frames_per_clip)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return nframes / dt, dt * 1e3, nframes


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    steps, warmup = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
    fps, ms, nframes = time_cpu(a, a.cpu_sample_frames, steps, warmup)
    cores = os.cpu_count() or 1
    sample = (f"2 clips x {a.cpu_sample_frames} frames = {nframes} frames per step (bounded sample of the "
              f"{a.clips * a.frames}-frame workload), {steps} timed steps after {warmup} warm-up")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "device": f"cpu ({cpu_model()})"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ Sinkhorn iters/s
def sinkhorn_rates(dev, cpu=True):
    """BASELINE.json's second metric: Sinkhorn iterations per second (1 iteration = one row + one column
    normalisation pass, graph_matching.py:659-669) of the fused instance-norm + sinkhorn_rpm(20) + exp kernel on
    252x252 matrices: one problem (the reference's shape, latency-bound) and 512 independent problems (saturating),
    CUDA-event timed; beside it the CPU oracle port of the same loop on the host cores."""
    from graphecho_b200 import functional as GF
    out = {"metric": "Sinkhorn iters/s (instance-norm + sinkhorn_rpm, 20 iterations, 252x252, slack row/column)"}
    for tag, batch in (("single_problem", 1), ("batched_512", 512)):
        M = torch.randn(batch, 252, 252, device=dev)
        for _ in range(3):
            GF.sinkhorn_rpm_exp(M, 20, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            GF.sinkhorn_rpm_exp(M, 20, True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[tag] = {"iters_per_s": 20 * batch / (ms / 1e3), "ms_per_call": ms}
    if cpu:
        from oracle import graph_ops as G
        Mc = torch.randn(252, 252)
        G.sinkhorn_rpm_exp(Mc, 20, True)
        t0 = time.perf_counter()
        for _ in range(10):
            G.sinkhorn_rpm_exp(Mc, 20, True)
        dt = (time.perf_counter() - t0) / 10
        out["cpu_port"] = {"iters_per_s": 20 / dt, "ms_per_call": dt * 1e3, "cores": os.cpu_count() or 1}
    return out


def ncu_traffic(entry):
    """DRAM bytes per call of a C-ABI entry point from the committed `ncu --set full` capture
    (profiles/r1d_ncu_kernels.json: every kernel of the entry at the [256,256,28,28] bf16 shape), or None."""
    fam = {"ge_bn_bwd": ("bn_partial_bwd", "bn_finalize_bwd", "bn_apply_bwd"),
           "ge_bn_fwd_train": ("bn_partial_stats", "bn_finalize_stats", "bn_apply_fwd"),
           "ge_gn_relu_upsample_bwd": ("gn_relu_up_bwd_reduce", "gn_relu_up_bwd_apply"),
           "ge_gn_relu_upsample_fwd": ("gn_relu_up_fwd",), "ge_group_stats": ("group_stats",),
           "ge_knn_graph_nmajor": ("knn_split_nmajor", "knn_tc_kernel"),
           "ge_mrconv_gather_nmajor_fwd": ("mr_gather_nmajor_fwd",),
           "ge_mrconv_gather_nmajor_bwd": ("mr_gather_nmajor_bwd_init", "mr_gather_nmajor_bwd_scatter")}.get(entry)
    try:
        rows = json.loads((ROOT / "profiles" / "r1d_ncu_kernels.json").read_text())
    except Exception:
        return None
    if not fam:
        return None
    tot = sum((r["rd"] + r["wr"]) * 1e6 for r in rows if any(r["kernel"].startswith(f) for f in fam))
    return tot or None


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch.distributed as dist
    from graphecho_b200 import _cabi
    from graphecho_b200.engine import EngineConfig, UDAEngine, init_distributed, make_batch, split_streams

    rank, local, world = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: graphecho_b200 has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True
    graphs = not a.no_graphs and not (a.sync_bn and world > 1)
    cfg = EngineConfig(hw=a.hw, num_classes=a.num_classes, bf16=not a.fp32, sync_bn=a.sync_bn,
                       cluster_backend="device", cuda_graphs=graphs)
    eng = UDAEngine(cfg, dev, world)
    clips_h, masks_h = make_batch(cfg, a.clips, a.frames, rank=rank, world=world, pin=True)
    clips_d, masks_d = clips_h.to(dev), masks_h.to(dev)
    frames_per_step = a.clips * a.frames

    def step_resident():
        fs, ft, shape = split_streams(clips_d)
        return eng.train_step(fs, masks_d, ft, shape)[0]

    def step_e2e():
        clips_d.copy_(clips_h, non_blocking=True)
        masks_d.copy_(masks_h, non_blocking=True)
        fs, ft, shape = split_streams(clips_d)
        return float(eng.train_step(fs, masks_d, ft, shape)[0])      # D2H read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            out = fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, out

    for _ in range(max(a.warmup, 3)):
        step_resident()
    launches0 = _cabi.launch_count()
    # one poller per job (rank 0): eight nvidia-smi loops contend for the driver and stretch every rank's
    # launch-bound graph-module section
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        ms, last = timed(step_resident, a.steps)
    launches = (_cabi.launch_count() - launches0) // max(a.steps, 1) + eng.graph_launches
    clocks = clk.summary()
    value = world * frames_per_step / (ms / 1e3)

    e2e = None
    if not a.skip_e2e:
        for _ in range(2):
            step_e2e()
        ms_e2e, _ = timed(step_e2e, a.steps)
        e2e = {"value": world * frames_per_step / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": clips_h.numel() * 4 + masks_h.numel() * 4, "d2h_bytes_per_step": 4}

    # per-kernel timing of the custom kernels (CUDA events on the launching stream): one extra step on
    # an eager (graph-free) twin of the engine, so that every entry point is individually visible
    if cfg.cuda_graphs:
        import copy
        ecfg = copy.copy(cfg)
        ecfg.cuda_graphs = False
        peng = UDAEngine(ecfg, dev, world)
    else:
        peng = eng

    def step_profile():
        fs, ft, shape = split_streams(clips_d)
        return peng.train_step(fs, masks_d, ft, shape)[0]

    for _ in range(2 if peng is not eng else 0):
        step_profile()
    _cabi.profile_start()
    step_profile()
    prof = _cabi.profile_stop()
    del peng
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    kernels = {}
    for name, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        per = r["ms"] / r["calls"]
        kernels[name] = {"calls": r["calls"], "ms": round(r["ms"], 4), "share_of_step": round(r["ms"] / ms, 4),
                         "GBps": round(r["bytes"] / max(r["ms"], 1e-9) / 1e6, 1),
                         "GFLOPs": round(r["flops"] / max(r["ms"], 1e-9) / 1e6, 1), "ms_per_call": round(per, 5)}
    # dominant HBM-bound custom kernel (entry points whose roofline is memory: everything except the
    # FFMA-bound k-NN / affinity kernels, which are reported in `kernels` with their GFLOP/s)
    hbm_names = [n for n in kernels if n not in ("ge_knn_graph", "ge_affinity_pairwise_fwd", "ge_affinity_pairwise_bwd")]
    roofline = None
    if hbm_names:
        top = hbm_names[0]
        r = prof[top]
        ach = r["bytes"] / r["ms"] / 1e6
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": ach / hbm_peak, "traffic": ncu_traffic(top),
                    "traffic_note": "dram__bytes_read+write of one call at the largest map of the step ([256,256,28,28] bf16, "
                                    "103 MB), ncu --set full, profiles/r1d_ncu_kernels.json; `achieved` averages all "
                                    "launches of the step (most maps are smaller and launch-bound)",
                    "launches_per_step": r["calls"], "avg_ms": r["ms"] / r["calls"]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        import warnings
        warnings.filterwarnings("ignore")
        fps, cms, nframes = time_cpu(a, a.cpu_sample_frames, 2, 1)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"2 clips x {a.cpu_sample_frames} frames = {nframes} frames/step, 2 timed steps "
                                  f"after 1 warm-up, {cms:.0f} ms/step, {cpu_model()}"}

    sinkhorn = sinkhorn_rates(dev, cpu=not a.skip_cpu_baseline) if rank == 0 else None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if a.fp32 else "bf16", "data": "synthetic",
                "config": {"workload": workload_name(a), "global_frames_per_step": world * frames_per_step,
                           "parallelism": f"dp{world}", "sync_bn": bool(cfg.sync_bn and world > 1), "cuda_graphs": bool(cfg.cuda_graphs),
                           "l2": "per-step working set (activations > 4 GB) exceeds the 126 MB L2; no flush needed",
                           "grad_allreduce_bytes": eng.grads.nbytes, "loss": float(last)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "sinkhorn": sinkhorn, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
