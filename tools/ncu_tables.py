#!/usr/bin/env python
"""Markdown tables of the per-kernel ncu summaries kept under profiles/ (tools/ncu_extract.py output).
  python tools/ncu_tables.py profiles/r2_ncu_full_fem128.csv ...
"""
import csv
import sys


def table(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    out = ["| kernel | grid × block | regs | time | DRAM read | DRAM write | DRAM % of peak | warps active | issue active "
           "| warp instr | top stalls (cycles per issue) |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    rd = wr = 0.0
    units = rows[1]

    def scale(col, v):  # bytes in GB whatever unit ncu chose
        u = units[ix[col]].lower()
        return v * {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0}.get(u, 1.0)

    def ms(col, v):
        u = units[ix[col]].lower()
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)

    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        g = lambda c: float(r[ix[c]].replace(",", "") or 0)
        r_gb, w_gb = scale("dram__bytes_read.sum", g("dram__bytes_read.sum")), scale("dram__bytes_write.sum", g("dram__bytes_write.sum"))
        stalls = sorted(((float(r[i] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for h, i in ix.items() if h.startswith("smsp__average_warps_issue_stalled_")), reverse=True)[:3]
        if "emit" not in name and "pack" not in name:
            rd += r_gb
            wr += w_gb
        out.append(f"| `{name}` | {int(g('launch__grid_size'))} × {int(g('launch__block_size'))} | {int(g('launch__registers_per_thread'))} "
                   f"| {ms('gpu__time_duration.sum', g('gpu__time_duration.sum')):.3f} ms | {r_gb:.3f} GB | {w_gb:.3f} GB "
                   f"| {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} % "
                   f"| {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} % "
                   f"| {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} % | {g('smsp__inst_executed.sum') / 1e6:.0f} M "
                   f"| {', '.join(f'{n} {v:.1f}' for v, n in stalls)} |")
    out.append("")
    out.append(f"DRAM bytes of one flush (all kernels but the producer): {rd:.3f} GB read + {wr:.3f} GB written = {rd + wr:.3f} GB.")
    return "\n".join(out)


for p in sys.argv[1:]:
    print(f"### {p}\n")
    print(table(p))
    print()
