"""profiles/r2_ncu_k3_k4.md from `ncu -i gpurun_out/k34_r2.ncu-rep --page raw --csv > raw.csv` (capture of scripts/ncu_k34.py)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "ms"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__cluster_dim_x", "cluster"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %")]
want = [(k, n) for k, n in want if k in col]
out = ["| # | kernel | " + " | ".join(n for _, n in want) + " |", "|---|---|" + "---:|" * len(want)]
for r in data:
    name = r[col["Kernel Name"]]
    if not any(t in name for t in ("affinity", "rpm", "sinkhorn")):
        continue
    short = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    vals = []
    for k, _ in want:
        v = r[col[k]]
        try:
            f = float(v.replace(",", ""))
            u = units[col[k]]
            if k.startswith("dram__bytes"):
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            if k.startswith("gpu__time"):
                f *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
            vals.append(f"{f:.4g}")
        except ValueError:
            vals.append(v)
    out.append(f"| {r[col['ID']]} | `{short}` | " + " | ".join(vals) + " |")
print("\n".join(out))
