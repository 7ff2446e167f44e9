"""Turn `ncu -i X.ncu-rep --page raw --csv` (or a `--metrics ... --csv` launch list) into the committed summaries.

    python scripts/ncu_report.py kernels gpurun_out/kernels_r2.csv profiles/r2_ncu_kernels      # -> .json + .md
    python scripts/ncu_report.py launches gpurun_out/launches_r2.csv profiles/r2_launches_step.md [step_ms]
"""
import csv, json, re, sys
from collections import OrderedDict, defaultdict

OURS = re.compile(r"bn_|gn_|group_stats|knn_|mr_gather|affinity|sinkhorn|seg_tail|seg_loss|mask_boxes|tgcn_|conv1x1_tc|maxpool3s2|"
                  r"upsample_|chan_stats|spectral|sd_|pool_concat")


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("ge_bn_coop::", "")
    return name.strip()


def load(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 5]
    hdr = rows[0]
    return hdr, rows[1:]


def kernels(path, out):
    hdr, rows = load(path)
    if "Metric Name" in hdr:            # long format
        ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        d = OrderedDict()
        for r in rows:
            d.setdefault((r[ii], r[ki]), {})[r[mi]] = r[vi]
        items = [(k[1], m) for k, m in d.items()]
    else:                               # wide (page raw): second row = units; values are normalised to ms / MB here
        ki = hdr.index("Kernel Name")
        units = dict(zip(hdr, rows[0]))
        tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
        bscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        items = []
        for r in rows[1:]:
            m = dict(zip(hdr, r))
            for k in list(m):
                u = units.get(k, "")
                try:
                    if u in tscale and "time_duration" in k:
                        m[k] = str(float(m[k].replace(",", "")) * tscale[u] * 1e6)       # -> ns (scaled back below)
                    elif u in bscale and k.startswith("dram__bytes_"):
                        m[k] = str(float(m[k].replace(",", "")) * bscale[u] * 1e6)       # -> bytes
                except ValueError:
                    pass
            items.append((r[ki], m))
    def f(m, key, scale=1.0):
        v = m.get(key)
        if v is None:                   # page raw prefixes some metrics with their section (e.g. FBSP.TriageCompute.)
            v = next((m[k] for k in m if k.endswith("." + key) or k.endswith(key)), "0")
        try:
            return float(str(v).replace(",", "") or 0) * scale
        except ValueError:
            return 0.0
    res = []
    for name, m in items:
        res.append({"kernel": short(name), "ms": f(m, "gpu__time_duration.sum", 1e-6),
                    "rd": f(m, "dram__bytes_read.sum", 1e-6), "wr": f(m, "dram__bytes_write.sum", 1e-6),
                    "dram%": f(m, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "dramGBps": (f(m, "dram__bytes_read.sum", 1e-6) + f(m, "dram__bytes_write.sum", 1e-6)) / max(f(m, "gpu__time_duration.sum", 1e-6), 1e-9),
                    "sm%": f(m, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "tensor%": f(m, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active") or
                               f(m, "sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active"),
                    "occ%": f(m, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "regs": f(m, "launch__registers_per_thread"), "grid": f(m, "launch__grid_size"),
                    "block": f(m, "launch__block_size")})
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as fh:
        fh.write("| kernel | ms | DRAM rd MB | DRAM wr MB | DRAM GB/s | SM % | tensor % | occ % | regs | grid | block |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in res:
            fh.write(f"| `{r['kernel']}` | {r['ms']:.4f} | {r['rd']:.1f} | {r['wr']:.1f} | {r['dramGBps']:.0f} | {r['sm%']:.1f} | {r['tensor%']:.1f} | "
                     f"{r['occ%']:.1f} | {r['regs']:.0f} | {r['grid']:.0f} | {r['block']:.0f} |\n")
    print(f"{len(res)} kernels -> {out}.json / .md")


def launches(path, out, step_ms=None):
    hdr, rows = load(path)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[mi] != "gpu__time_duration.sum":
            continue
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3          # ns -> us
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    groups = defaultdict(lambda: [0, 0.0])
    for k, v in agg.items():
        g = ("graphecho_b200" if OURS.search(k) else "library conv/GEMM" if re.search(r"cutlass|cudnn|xmma|gemm|sgemm|nchwToNhwc|nhwcToNchw|cublas|gemv", k, re.I)
             else "cuSOLVER" if re.search(r"syev|rotate|jacobi", k) else "NCCL" if "nccl" in k.lower() else "ATen elementwise / reduce / copy / optimizer")
        groups[g][0] += v[0]
        groups[g][1] += v[1]
    with open(out, "w") as fh:
        fh.write(f"launches in the step: **{n}**, summed kernel time **{tot / 1e3:.1f} ms**" + (f" (step: {step_ms} ms with overlap across streams)" if step_ms else "") + "\n\n")
        fh.write("| group | launches | ms | share |\n|---|---:|---:|---:|\n")
        for g, v in sorted(groups.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {g} | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / tot:.1f}% |\n")
        fh.write("\n| us | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
            fh.write(f"| {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[0]} | `{k[:110]}` |\n")
    print(f"{n} launches, {tot / 1e3:.2f} ms -> {out}")


if __name__ == "__main__":
    if sys.argv[1] == "kernels":
        kernels(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
