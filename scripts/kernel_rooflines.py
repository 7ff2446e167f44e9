"""Per-kernel micro-benchmarks: every C-ABI entry point at the reference shape and at a saturating
problem count, CUDA-event timed on the launching stream, L2 flushed between iterations, reported as
achieved GB/s (algorithmic bytes) and GFLOP/s next to the measured HBM peak (MEASURED_PEAKS.json) and
the fp32 FFMA peak (SMs x 128 lanes x 2 x SM clock).  Output: one JSON object per line + a markdown table.

    python scripts/kernel_rooflines.py [--out profiles/rN_kernel_rooflines.md]
"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from graphecho_b200 import functional as GF, _cabi

dev = torch.device("cuda:0")
peaks = {}
try:
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
except Exception:
    pass
HBM = peaks.get("hbm_gbs", 6650.0)
FFMA = _cabi.lib().ge_device_sm_count() * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) / 1e3)   # GFLOP/s
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, iters=20, warm=3):
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
    ts.sort()
    return ts[len(ts) // 2]


rows = []


def report(name, shape, ms, nbytes, flops, bound):
    gbs, gf = nbytes / ms / 1e6, flops / ms / 1e6
    frac = gbs / HBM if bound == "hbm" else gf / FFMA
    rows.append(dict(kernel=name, shape=shape, ms=round(ms, 4), GBps=round(gbs, 1), GFLOPs=round(gf, 1),
                     bound=bound, frac=round(frac, 3)))
    print(json.dumps(rows[-1]), flush=True)


def bench_affinity():
    for batch, n in ((1, 252), (1, 320), (256, 252)):
        A, B = torch.randn(batch, n, 512, device=dev), torch.randn(batch, n, 512, device=dev)
        w2, b2 = torch.randn(512, device=dev), torch.randn(1, device=dev)
        ms = timed(lambda: GF.affinity_pairwise(A, B, w2, b2))
        report("ge_affinity_pairwise_fwd", f"{batch}x{n}x{n}x512", ms, 4 * batch * (512 * 2 * n + n * n), 3 * batch * 512 * n * n, "fp32")
        Ar, Br = A.clone().requires_grad_(), B.clone().requires_grad_()
        M = GF.affinity_pairwise(Ar, Br, w2, b2)
        g = torch.randn_like(M)
        ms = timed(lambda: torch.autograd.grad(M, (Ar, Br), g, retain_graph=True))
        report("ge_affinity_pairwise_bwd", f"{batch}x{n}x{n}x512", ms, 4 * batch * (512 * 4 * n + n * n), 8 * batch * 512 * n * n, "fp32")


def bench_sinkhorn():
    for batch, n in ((1, 252), (1, 320), (64, 252), (512, 252)):
        M = torch.randn(batch, n, n, device=dev)
        ms = timed(lambda: GF.sinkhorn_rpm_exp(M, 20, True))
        report("ge_sinkhorn_rpm_fwd(20 it)", f"{batch}x{n}x{n}", ms, 8 * batch * n * n, batch * n * n * 166, "hbm")
        rows[-1]["sinkhorn_iters_per_s"] = round(20 * batch / (ms / 1e3))
        print(json.dumps({"sinkhorn_iters_per_s": rows[-1]["sinkhorn_iters_per_s"], "shape": f"{batch}x{n}x{n}"}))
        Mr = M.clone().requires_grad_()
        P = GF.sinkhorn_rpm_exp(Mr, 20, True)
        g = torch.randn_like(P)
        ms = timed(lambda: torch.autograd.grad(P, Mr, g, retain_graph=True))
        report("ge_sinkhorn_rpm_bwd(20 it)", f"{batch}x{n}x{n}", ms, 12 * batch * n * n, batch * n * n * 250, "hbm")
    for B, P_ in ((4, 64), (16, 64), (256, 64)):
        x, y = torch.randn(B, P_, 256, device=dev) * 0.05, torch.randn(B, P_, 256, device=dev) * 0.05
        ms = timed(lambda: GF.sinkhorn_distance(x, y, 0.1, 5))
        report("ge_sinkhorn_distance_fwd(5 it)", f"{B}x{P_}x{P_}x256", ms, 4 * B * 256 * 2 * P_ + 8 * B * P_ * P_, 3 * B * P_ * P_ * 256, "hbm")
        rows[-1]["sinkhorn_iters_per_s"] = round(5 * B / (ms / 1e3))


TENSOR = peaks.get("bf16_tflops", 1590.0) * 1e3   # GFLOP/s, dense 16-bit tensor peak (cuBLAS burst)


def bench_knn():
    lib = _cabi.lib()
    for B, C, N, k in ((8, 256, 64, 9), (256, 256, 784, 9), (16, 256, 4096, 9)):
        x = torch.randn(B, C, N, 1, device=dev)
        for path, label in ((1, "ge_knn_graph[fp32 FFMA]"), (0, "ge_knn_graph[tcgen05 2xfp16-split]")):
            if path == 0 and N < 128:
                continue
            lib.ge_knn_graph_set_path(path)
            ms = timed(lambda: GF.knn_graph(x, None, k, 1), iters=10)
            lib.ge_knn_graph_set_path(0)
            flops = 2 * B * N * N * C
            if path == 1:
                report(label, f"B{B} C{C} N{N} k{k}", ms, 8 * B * C * N + 16 * B * N * k, flops, "fp32")
            else:   # tensor roofline: the split issues 3x the algorithmic flops on the 16-bit tensor pipe
                rows.append(dict(kernel=label, shape=f"B{B} C{C} N{N} k{k}", ms=round(ms, 4),
                                 GBps=round((8 * B * C * N + 16 * B * N * k) / ms / 1e6, 1), GFLOPs=round(flops / ms / 1e6, 1),
                                 bound="tensor(3x)", frac=round(3 * flops / ms / 1e6 / TENSOR, 3)))
                print(json.dumps(rows[-1]), flush=True)
        e = GF.knn_graph(x, None, k, 1)
        ms = timed(lambda: GF.mr_gather(x, e, None), iters=10)
        report("ge_mrconv_gather_fwd", f"B{B} C{C} N{N} k{k}", ms, 4 * B * C * N + 8 * B * N * k + 9 * B * C * N, 2 * B * C * N * k, "hbm")


def bench_maps():
    for dt in (torch.bfloat16, torch.float32):
        es = 2 if dt == torch.bfloat16 else 4
        N, C, H = 256, 256, 28
        x = torch.randn(N, C, H, H, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        bn = torch.nn.BatchNorm2d(C).to(dev)
        res = torch.randn_like(x)
        ms = timed(lambda: GF.bn_act(x, bn, residual=res, relu=True))
        report("ge_bn_fwd_train(+res+relu)", f"{N}x{C}x{H}x{H} {dt}", ms, N * C * H * H * es * 4, 8 * N * C * H * H, "hbm")
        xr = x.clone().requires_grad_()
        out = GF.bn_act(xr, bn, residual=None, relu=True)
        g = torch.randn_like(out)
        ms = timed(lambda: torch.autograd.grad(out, xr, g, retain_graph=True))
        report("ge_bn_bwd(relu)", f"{N}x{C}x{H}x{H} {dt}", ms, N * C * H * H * es * 4, 16 * N * C * H * H, "hbm")
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ms = timed(lambda: GF.gn_relu(x, gamma, beta, 32))
        report("ge_group_stats+gn_relu_fwd(G=32)", f"{N}x{C}x{H}x{H} {dt}", ms, N * C * H * H * es * 3, 15 * N * C * H * H, "hbm")
        out = GF.gn_relu(xr, gamma, beta, 32)
        ms = timed(lambda: torch.autograd.grad(out, xr, g, retain_graph=True))
        report("ge_gn_relu_upsample_bwd(G=32)", f"{N}x{C}x{H}x{H} {dt}", ms, N * C * H * H * es * 5, 20 * N * C * H * H, "hbm")
        top = torch.randn(N, C, 14, 14, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        ms = timed(lambda: GF.upsample_add(top, x))
        report("ge_upsample_add_fwd", f"{N}x{C} 14->28 {dt}", ms, N * C * es * (14 * 14 + 2 * H * H), 8 * N * C * H * H, "hbm")
        s = [torch.randn(N, 128, H, H, device=dev).to(dt).contiguous(memory_format=torch.channels_last) for _ in range(4)]
        W3, b3 = torch.randn(2, 128, 1, 1, device=dev), torch.randn(2, device=dev)
        ms = timed(lambda: GF.seg_tail(*s, W3, b3, 4))
        report("ge_seg_tail_fwd", f"{N}x128x{H}x{H} nc=2 {dt}", ms, 4 * N * H * H * 128 * es + 4 * N * 2 * 16 * H * H, 2 * N * H * H * 128 * 4, "hbm")
    lv = [torch.randn(64, 256, s_, s_, device=dev).contiguous(memory_format=torch.channels_last) for s_ in (64, 32, 16, 8)]
    ms = timed(lambda: GF.pool_concat(lv, (8, 4, 2, 1)))
    report("ge_tgcn_pool_concat_fwd", "64 frames @256^2 pyramid fp32", ms, 64 * 4 * 256 * (64 * 64 + 32 * 32 + 16 * 16 + 8 * 8) + 64 * 1024 * 64 * 4, 64 * 256 * 5440, "hbm")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    print(json.dumps({"hbm_peak_GBps": HBM, "ffma_peak_GFLOPs": round(FFMA), "source": "MEASURED_PEAKS.json" if peaks else "fallback"}))
    bench_affinity(); bench_sinkhorn(); bench_knn(); bench_maps()
    if a.out:
        with open(a.out, "w") as f:
            f.write(f"# Kernel micro-benchmarks (CUDA events, L2 flushed, median of 20)\\n\\nHBM peak {HBM} GB/s (measured), fp32 FFMA peak {FFMA:.0f} GFLOP/s (SMs x 128 x 2 x max SM clock).\\n"
                    "`frac` = achieved / peak of the bound that applies (hbm: algorithmic bytes; fp32: algorithmic flops).\\n\\n")
            f.write("| kernel | shape | ms | GB/s | GFLOP/s | bound | frac |\\n|---|---|---:|---:|---:|---|---:|\\n")
            for r in rows:
                f.write(f"| `{r['kernel']}` | {r['shape']} | {r['ms']} | {r['GBps']} | {r['GFLOPs']} | {r['bound']} | {r['frac']} |\\n")
