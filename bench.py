#!/usr/bin/env python
"""Benchmark of GraphEcho's data-parallel hot path (BASELINE.json metric: frames/s through one full UDA training
step -- forward + backward + optimizers -- on synthetic inputs of the named shapes).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5]     # ours (one rank per GPU under torchrun)
  python bench.py --impl reference ...    # the reference's own CPU implementation of the same step (oracle/_ref)

--config (BASELINE.json `configs`); default 2, the configuration the metric is quoted on:
  2  8 clips x 32 frames 112x112 per GPU, FPN(resnet, nc=2) + ViG Grapher(p2) + GModule + 4 discriminators, bf16
  3  16 source + 16 target frames 256x256, FPN(resnet, nc=4) + GModule + discriminators + SinkhornDistance on the node sets
  4  16 + 8 single frames + 8 clips x 8 frames 256x256 per GPU, FPN(VGG16, nc=3) + GModule + discriminators + TGCN
  5  config 2 at 16 clips per GPU (the 8-GPU capture run)

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

METRIC = "frames/s through one full UDA training step (forward + backward + optimizers)"
UNIT = "frames/s"

# per-GPU workload of every config: (clips, frames per clip) for the clip configs, (source, target) single frames and
# (clips, frames) of the temporal branch for the CardiacUDA-shaped ones
WORK = {2: dict(clips=8, frames=32), 5: dict(clips=16, frames=32),
        3: dict(n_src=16, n_tgt=16), 4: dict(n_src=16, n_tgt=8, clips=8, frames=8)}
# bounded CPU samples of the same workloads (about 10-30 s of CPU work for 2 warm-up + 5 timed steps)
CPU_SAMPLE = {2: dict(clips=2, frames=16), 5: dict(clips=2, frames=16),
              3: dict(n_src=2, n_tgt=2), 4: dict(n_src=2, n_tgt=1, clips=2, frames=2)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU (half source, half target); default per config")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--sync-bn", action="store_true",
                    help="SyncBatchNorm on the network when N>1 (the reference's intent): fused BatchNorm kernels with the two per-layer NCCL averages; eager, no CUDA graphs)")
    ap.add_argument("--no-graphs", action="store_true", help="run the static segments eagerly instead of as CUDA graphs")
    ap.add_argument("--fp32", action="store_true", help="fp32 convolutions instead of bf16 autocast")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-rooflines", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="--impl reference: time the oracle port even when oracle/_ref exists")
    a = ap.parse_args()
    w = dict(WORK[a.config])
    if a.clips is not None:
        w["clips"] = a.clips
    if a.frames is not None:
        w["frames"] = a.frames
    a.work = w
    return a


def frames_per_step(w):
    return w.get("n_src", 0) + w.get("n_tgt", 0) + w.get("clips", 0) * w.get("frames", 0)


def workload_name(cfg_id, w):
    if cfg_id in (2, 5):
        return (f"config{cfg_id}: {w['clips']} clips x {w['frames']} frames 112x112 per GPU (half source, half target), "
                f"FPN(resnet,nc=2) + ViG Grapher(p2,k=9) + GModule + 4 Discriminators, fwd+bwd+optim")
    if cfg_id == 3:
        return (f"config3: {w['n_src']} source + {w['n_tgt']} target frames 256x256 per GPU, FPN(resnet,nc=4) + GModule + "
                f"4 Discriminators + SinkhornDistance(0.1,5,'mean') on the node sets, fwd+bwd+optim")
    return (f"config4: {w['n_src']} source + {w['n_tgt']} target frames + {w['clips']} clips x {w['frames']} frames 256x256 per GPU, "
            f"FPN(VGG16,nc=3) + GModule + 4 Discriminators + TGCN temporal module, fwd+bwd+optim")


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


# ------------------------------------------------------------------------------------------ inputs
def host_inputs(cfg, w, rank=0, pin=False):
    """Synthetic step input on the host: dict with frames_src / masks_src / frames_tgt (single-frame configs) or
    clips / masks (clip configs), plus the temporal clips of config 4."""
    from graphecho_b200.engine import make_batch, make_frame_batch
    out = {}
    if "n_src" in w:
        out["xs"], out["masks"], out["xt"] = make_frame_batch(cfg, w["n_src"], w["n_tgt"], rank=rank, pin=pin)
        if "clips" in w:
            out["tclips"], out["tmasks"] = make_batch(cfg, w["clips"], w["frames"], rank=rank, pin=pin)
    else:
        out["clips"], out["masks"] = make_batch(cfg, w["clips"], w["frames"], rank=rank, pin=pin)
    return out


def step_args(dev_in):
    """(frames_src, masks_src, frames_tgt, temporal) from device-resident inputs."""
    from graphecho_b200.engine import split_streams, temporal_input
    if "clips" in dev_in:
        fs, ft, _ = split_streams(dev_in["clips"])
        return fs, dev_in["masks"], ft, None
    temporal = temporal_input(dev_in["tclips"], dev_in["tmasks"]) if "tclips" in dev_in else None
    return dev_in["xs"], dev_in["masks"], dev_in["xt"], temporal


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_step_runner(cfg_id, w, use_reference=True):
    """One step of the same workload on the host cores: the reference's own modules (oracle/_ref, kind 'reference')
    when staged, else the oracle port (kind 'port')."""
    from graphecho_b200 import synth
    from graphecho_b200.engine import preset
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = preset(cfg_id, clip_frames=w.get("frames", 8)) if cfg_id == 4 else preset(cfg_id)
    h = host_inputs(cfg, w)
    if "clips" in h:
        frames = synth.flatten_clips(h["clips"])
        ns = frames.shape[0] // 2
        fs, masks, ft, temporal = frames[:ns], h["masks"], frames[ns:], None
    else:
        fs, masks, ft = h["xs"], h["masks"], h["xt"]
        temporal = (synth.flatten_clips(h["tclips"]), h["tmasks"], (w["clips"], w["frames"])) if "tclips" in h else None
    nc, bb = cfg.num_classes, cfg.backbone
    kw = dict(sinkhorn_nodes=cfg.sinkhorn_nodes, sinkhorn_weight=cfg.sinkhorn_weight, temporal=temporal)
    kind = "port"
    if use_reference:
        try:
            from oracle import ref_step as RS
            if RS.reference_root() is not None:
                M = RS.build_modules(nc, bb, grapher=cfg.vig_grapher, tgcn=cfg.temporal_graph,
                                     clip_frames=w.get("frames", 8), hw=cfg.hw)
                opt = RS.build_optimizers(M)
                kind = "reference"

                def step():
                    return float(RS.train_step(M, opt, fs, masks, ft, **kw)[0])
        except Exception as e:                                       # pragma: no cover - reported in the line
            sys.stderr.write(f"reference arm unavailable ({e!r}); timing the oracle port\n")
            kind = "port"
    if kind == "port":
        from oracle import step as OS
        P = OS.build_params(nc, bb, grapher=cfg.vig_grapher, tgcn=cfg.temporal_graph, clip_frames=w.get("frames", 8))
        opt = OS.build_optimizers(P)

        def step():
            return float(OS.train_step(P, opt, fs, masks, ft, num_classes=nc, backbone=bb, **kw)[0])

    return step, frames_per_step(w), kind


def time_cpu(cfg_id, w, steps, warmup, budget_s=150.0, use_reference=True):
    """-> (frames/s from the MEDIAN step, median ms, frames per step, kind, timed steps)."""
    import warnings
    warnings.filterwarnings("ignore")
    step, nframes, kind = cpu_step_runner(cfg_id, w, use_reference)
    for _ in range(warmup):
        step()
    ts, t_start = [], time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(ts) >= 5:
            break
    med = statistics.median(ts)
    return nframes / med, med * 1e3, nframes, kind, len(ts)


def sample_text(cfg_id, ws, w, steps, warmup, ms):
    parts = []
    if "n_src" in ws:
        parts.append(f"{ws['n_src']} source + {ws['n_tgt']} target frames")
    if "clips" in ws:
        parts.append(f"{ws['clips']} clips x {ws['frames']} frames")
    return (f"config{cfg_id} modules on {' + '.join(parts)} = {frames_per_step(ws)} frames per step (bounded sample of the "
            f"{frames_per_step(w)}-frame workload), median of {steps} timed steps after {warmup} warm-up, {ms:.0f} ms/step, "
            f"fp32, {cpu_model()}")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ws = CPU_SAMPLE[a.config]
    steps, warmup = max(5, a.steps), max(2, min(a.warmup, 5))
    fps, ms, nframes, kind, done = time_cpu(a.config, ws, steps, warmup, use_reference=not a.cpu_port)
    cores = os.cpu_count() or 1
    sample = sample_text(a.config, ws, a.work, done, warmup, ms)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": done,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.config, a.work), "device": f"cpu ({cpu_model()})",
                       "implementation": "unmodified reference modules (oracle/_ref) in the reference trainer's step order"
                       if kind == "reference" else "oracle port (oracle/step.py)"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ north-star kernel rooflines
def _timed_kernel(fn, flush, iters=15, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def north_star_rooflines(dev, peaks):
    """The kernels BASELINE.json's north star names, each alone (CUDA events, L2 flushed between iterations, median of
    15): algorithmic bytes / flops per SURVEY.md section 8(d) over the measured time, against the measured HBM peak,
    the fp32 FFMA peak (SMs x 128 lanes x 2 x max SM clock) or the measured bf16 tensor peak.  Reference shape (one
    problem, latency-bound) and a saturating count, as section 8(d) asks."""
    from graphecho_b200 import functional as GF, _cabi
    hbm = peaks.get("hbm_gbs") or 6650.0
    ffma = _cabi.lib().ge_device_sm_count() * 128 * 2 * ((peaks.get("sm_max_mhz") or 1965.0) / 1e3)       # GFLOP/s
    tensor = (peaks.get("bf16_tflops") or 1590.0) * 1e3
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []

    def add(kernel, shape, ms, nbytes, flops, bound, extra=None):
        gbs, gf = nbytes / ms / 1e6, flops / ms / 1e6
        ach, peak, unit = {"hbm": (gbs, hbm, "GB/s"), "fp32": (gf, ffma, "GFLOP/s"), "tensor": (gf, tensor, "GFLOP/s")}[bound]
        r = {"kernel": kernel, "shape": shape, "ms": round(ms, 4), "bound": bound, "achieved": round(ach, 1),
             "peak": round(peak, 1), "unit": unit, "frac": round(ach / peak, 4), "GBps": round(gbs, 1), "GFLOPs": round(gf, 1)}
        r.update(extra or {})
        rows.append(r)

    w2, b2 = torch.randn(512, device=dev), torch.randn(1, device=dev)
    for batch, n in ((1, 252), (512, 252)):
        A, B = torch.randn(batch, n, 512, device=dev), torch.randn(batch, n, 512, device=dev)
        # K3 is instruction-issue bound (no GEMM in it: the ReLU couples i, j and k).  "fp32" rows count fp32 pipe
        # operations against SMs x 128 lanes x 2 x clock, i.e. an FFMA counts 2 and a max / gated add counts 1:
        # forward = max + fma per term (3 ops, 2 issue slots), backward = gate + 2 gated adds per term (3 ops, 3 slots).
        # `issue_frac` is the same time against the issue-slot bound itself (SMs x 128 lanes x clock).
        issue = ffma / 2
        ms = _timed_kernel(lambda: GF.affinity_pairwise(A, B, w2, b2), flush)
        add("ge_affinity_pairwise_fwd", f"{batch}x{n}x{n}x512", ms, 4 * batch * (512 * 2 * n + n * n), 3 * batch * 512 * n * n, "fp32",
            {"issue_frac": round(2 * batch * 512 * n * n / ms / 1e6 / issue, 4)})
        Ar, Br = A.clone().requires_grad_(), B.clone().requires_grad_()
        Mx = GF.affinity_pairwise(Ar, Br, w2, b2)
        g = torch.randn_like(Mx)
        ms = _timed_kernel(lambda: torch.autograd.grad(Mx, (Ar, Br), g, retain_graph=True), flush)
        add("ge_affinity_pairwise_bwd", f"{batch}x{n}x{n}x512", ms, 4 * batch * (512 * 4 * n + 2 * n * n), 3 * batch * 512 * n * n, "fp32",
            {"issue_frac": round(3 * batch * 512 * n * n / ms / 1e6 / issue, 4)})
        M = torch.randn(batch, n, n, device=dev)
        ms = _timed_kernel(lambda: GF.sinkhorn_rpm_exp(M, 20, True), flush)
        # 8 N1 N2 bytes for the whole loop (section 8(d) K4).  The loop itself is on-chip; in the exponent domain it is
        # 2 matrix-vector products per iteration = 4 N1 N2 flops per iteration on the fp32 pipe (`fp32_frac`); what
        # actually bounds it is the per-iteration dependency chain (reduce -> exchange -> reciprocal), see DESIGN.md
        add("ge_sinkhorn_rpm_fwd(instnorm+20it+exp)", f"{batch}x{n}x{n}", ms, 8 * batch * n * n, batch * n * n * (4 * 20 + 12), "hbm",
            {"sinkhorn_iters_per_s": round(20 * batch / (ms / 1e3)),
             "fp32_frac": round(batch * n * n * (4 * 20 + 12) / ms / 1e6 / ffma, 4)})
        Mr = M.clone().requires_grad_()
        Pm = GF.sinkhorn_rpm_exp(Mr, 20, True)
        g = torch.randn_like(Pm)
        ms = _timed_kernel(lambda: torch.autograd.grad(Pm, Mr, g, retain_graph=True), flush)
        add("ge_sinkhorn_rpm_bwd", f"{batch}x{n}x{n}", ms, 12 * batch * n * n, batch * n * n * (8 * 20 + 16), "hbm",
            {"fp32_frac": round(batch * n * n * (8 * 20 + 16) / ms / 1e6 / ffma, 4)})
    Bf, C, N, k = 256, 256, 784, 9
    xn = torch.randn(Bf, N, C, device=dev).bfloat16()
    ms = _timed_kernel(lambda: GF.knn_graph_nmajor(xn, None, k, 1), flush, iters=10)
    add("ge_knn_graph_nmajor[tcgen05]", f"B{Bf} N{N} C{C} k{k} bf16", ms, 2 * Bf * C * 2 * N + 16 * Bf * N * k, 2 * Bf * N * N * C, "tensor",
        {"tensor_frac_issued_3x": round(3 * 2 * Bf * N * N * C / ms / 1e6 / tensor, 4)})
    e = GF.knn_graph_nmajor(xn, None, k, 1)[0]
    ms = _timed_kernel(lambda: GF.mr_gather_nmajor(xn, e), flush, iters=10)
    add("ge_mrconv_gather_nmajor_fwd", f"B{Bf} N{N} C{C} k{k} bf16", ms, 2 * Bf * C * N + 8 * Bf * N * k + (2 * 2 + 1) * Bf * C * N, 2 * Bf * C * N * k, "hbm")
    xr = xn.clone().requires_grad_()
    f = GF.mr_gather_nmajor(xr, e)
    g = torch.randn_like(f)
    ms = _timed_kernel(lambda: torch.autograd.grad(f, xr, g, retain_graph=True), flush, iters=10)
    add("ge_mrconv_gather_nmajor_bwd", f"B{Bf} N{N} C{C} k{k} bf16", ms, (2 * 2 + 1 + 2) * Bf * C * N + 8 * Bf * N * k, 2 * Bf * C * N, "hbm")
    # the dominant glue kernels at the largest map of the step
    cl = torch.channels_last
    x = torch.randn(256, 256, 28, 28, device=dev).bfloat16().contiguous(memory_format=cl)
    res = torch.randn_like(x)
    bn = torch.nn.BatchNorm2d(256).to(dev)
    nel = x.numel()
    ms = _timed_kernel(lambda: GF.bn_act(x, bn, residual=res, relu=True), flush)
    add("ge_bn_fwd_train(+res+relu)", "256x256x28x28 bf16", ms, nel * 2 * 3, 8 * nel, "hbm")
    xr = x.clone().requires_grad_()
    out = GF.bn_act(xr, bn, residual=None, relu=True)
    g = torch.randn_like(out)
    ms = _timed_kernel(lambda: torch.autograd.grad(out, xr, g, retain_graph=True), flush)
    add("ge_bn_bwd(relu)", "256x256x28x28 bf16", ms, nel * 2 * 3 + nel // 8, 16 * nel, "hbm")
    return rows


def ncu_traffic(entry):
    """DRAM bytes per call of a C-ABI entry point from the committed `ncu --set full` capture of the CURRENT kernels
    (profiles/r2_ncu_kernels.json: every kernel of the entry at the [256,256,28,28] bf16 shape), or None."""
    fam = {"ge_bn_bwd": ("bn_partial_bwd", "bn_finalize_bwd", "bn_apply_bwd4"),
           "ge_bn_fwd_train": ("bn_partial_stats", "bn_finalize_stats", "bn_apply_fwd4"),
           "ge_gn_relu_upsample_bwd": ("gn_relu_bwd", "gn_relu_up_bwd"),
           "ge_gn_relu_upsample_fwd": ("gn_relu_up_fwd", "gn_relu_fwd"), "ge_group_stats": ("group_stats",),
           "ge_group_stats_bias": ("group_stats_bias",),
           "ge_knn_graph_nmajor": ("knn_split_nmajor", "knn_tc_kernel"),
           "ge_mrconv_gather_nmajor_fwd": ("mr_gather_nmajor_fwd",),
           "ge_mrconv_gather_nmajor_bwd": ("mr_gather_nmajor_bwd_init", "mr_gather_nmajor_bwd_scatter")}.get(entry)
    try:
        rows = json.loads((ROOT / "profiles" / "r2_ncu_kernels.json").read_text())
    except Exception:
        return None
    if not fam:
        return None
    tot = sum((r["rd"] + r["wr"]) * 1e6 for r in rows if any(r["kernel"].startswith(f) for f in fam))
    return tot or None


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import copy
    import torch.distributed as dist
    from graphecho_b200 import _cabi
    from graphecho_b200.engine import UDAEngine, init_distributed, preset

    rank, local, world = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: graphecho_b200 has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True
    w = a.work
    graphs = not a.no_graphs and not (a.sync_bn and world > 1)
    over = dict(bf16=not a.fp32, sync_bn=a.sync_bn, cluster_backend="device", cuda_graphs=graphs)
    if a.config == 4:
        over["clip_frames"] = w["frames"]
    cfg = preset(a.config, **over)
    eng = UDAEngine(cfg, dev, world)
    host = host_inputs(cfg, w, rank=rank, pin=True)
    devin = {k: v.to(dev) for k, v in host.items()}
    fps_frames = frames_per_step(w)
    grad_bytes = eng.grads.nbytes

    def step_resident():
        return eng.train_step(*step_args(devin))[0]

    def step_e2e():
        for k, v in host.items():
            devin[k].copy_(v, non_blocking=True)
        return float(eng.train_step(*step_args(devin))[0])      # D2H read of the step's loss

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

    warm = max(a.warmup, 3)
    for _ in range(warm):
        step_resident()
    launches0 = _cabi.launch_count()
    # one poller per job (rank 0): eight nvidia-smi loops contend for the driver and stretch every rank's
    # launch-bound graph-module section
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        ms, last = timed(step_resident, a.steps)
    launches = (_cabi.launch_count() - launches0) // max(a.steps, 1) + eng.graph_launches
    clocks = clk.summary()
    value = world * fps_frames / (ms / 1e3)

    e2e = None
    if not a.skip_e2e:
        for _ in range(2):
            step_e2e()
        ms_e2e, _ = timed(step_e2e, a.steps)
        e2e = {"value": world * fps_frames / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()), "d2h_bytes_per_step": 4}

    # per-kernel timing of the custom kernels (CUDA events on the launching stream): one extra step on
    # an eager (graph-free) twin of the engine, so that every entry point is individually visible
    if cfg.cuda_graphs:
        ecfg = copy.copy(cfg)
        ecfg.cuda_graphs = False
        peng = UDAEngine(ecfg, dev, world)
    else:
        peng = eng

    def step_profile():
        return peng.train_step(*step_args(devin))[0]

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
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "MEASURED_PEAKS.json") if peaks.get("hbm_gbs") else (6650.0, "B200_PROFILING.md fallback")
    kernels = {}
    for name, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        per = r["ms"] / r["calls"]
        kernels[name] = {"calls": r["calls"], "ms": round(r["ms"], 4), "share_of_step": round(r["ms"] / ms, 4),
                         "GBps": round(r["bytes"] / max(r["ms"], 1e-9) / 1e6, 1),
                         "GFLOPs": round(r["flops"] / max(r["ms"], 1e-9) / 1e6, 1), "ms_per_call": round(per, 5)}
    # dominant HBM-bound custom kernel: `achieved` = COMPULSORY bytes of all its launches in the step (every operand
    # read once, every result written once -- the `work=` figures in graphecho_b200/functional.py) / their summed time
    hbm_names = [n for n in kernels if n not in ("ge_knn_graph", "ge_knn_graph_nmajor", "ge_affinity_pairwise_fwd",
                                                 "ge_affinity_pairwise_bwd", "ge_sinkhorn_rpm_fwd", "ge_sinkhorn_rpm_bwd",
                                                 "ge_sinkhorn_distance_fwd", "ge_sinkhorn_distance_bwd")]
    roofline = None
    if hbm_names:
        top = hbm_names[0]
        r = prof[top]
        ach = r["bytes"] / r["ms"] / 1e6
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": ach / hbm_peak, "traffic": ncu_traffic(top),
                    "bytes_definition": "compulsory traffic: each operand read once, each result written once (functional.py work=)",
                    "traffic_note": "dram__bytes_read+write of one call at the largest map of the step ([256,256,28,28] bf16, "
                                    "103 MB), ncu --set full, profiles/r2_ncu_kernels.json; `achieved` averages all "
                                    "launches of the step (most maps are smaller and launch-bound)",
                    "launches_per_step": r["calls"], "avg_ms": r["ms"] / r["calls"]}

    rooflines = None
    if rank == 0 and world == 1 and not a.skip_rooflines:
        del eng
        torch.cuda.empty_cache()
        try:
            rooflines = north_star_rooflines(dev, peaks)
        except Exception as e:                                   # pragma: no cover
            rooflines = {"error": repr(e)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        ws = CPU_SAMPLE[a.config]
        fps, cms, nframes, kind, done = time_cpu(a.config, ws, 5, 2)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                        "sample": sample_text(a.config, ws, w, done, 2, cms)}

    if rank == 0:
        sink = None
        if rooflines and isinstance(rooflines, list):
            sk = [r for r in rooflines if r["kernel"].startswith("ge_sinkhorn_rpm_fwd")]
            sink = {"metric": "Sinkhorn iters/s (instance-norm + sinkhorn_rpm, 20 iterations, 252x252, slack row/column)",
                    "single_problem": sk[0]["sinkhorn_iters_per_s"], "batched_512": sk[1]["sinkhorn_iters_per_s"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if a.fp32 else "bf16", "data": "synthetic",
                "config": {"workload": workload_name(a.config, w), "config_id": a.config,
                           "global_frames_per_step": world * fps_frames,
                           "parallelism": f"dp{world}", "sync_bn": bool(cfg.sync_bn and world > 1), "cuda_graphs": bool(cfg.cuda_graphs),
                           "per_domain_bn": bool(cfg.per_domain_bn),
                           "l2": "per-step working set (activations > 4 GB) exceeds the 126 MB L2; no flush needed",
                           "grad_allreduce_bytes": grad_bytes, "loss": float(last)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "rooflines": rooflines,
                "cpu_baseline": cpu_baseline, "sinkhorn": sink, "kernels": kernels}
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
