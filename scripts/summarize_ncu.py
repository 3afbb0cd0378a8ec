#!/usr/bin/env python3
"""Turn the ncu outputs of scripts/gpu_r2_evidence.sh (launch list CSV + `--page raw --csv` of the full capture)
into the markdown summary kept under profiles/.  Usage:
    python scripts/summarize_ncu.py gpurun_out/launches_c4.csv gpurun_out/prof_sk_raw.csv > profiles/r02/ncu_summary_final.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0][:70]


def launch_table(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        k = short(r[4])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + float(r[-1]) / 1e3)
    total = sum(t for _, t in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / total >= 0.003:
            out.append("| %s | %d | %.1f | %.1f%% |" % (k, n, t, 100 * t / total))
    return "\n".join(out), total


FIELDS = [("time ms", "gpu__time_duration.sum", 1.0),
          ("dram read GB", "dram__bytes_read.sum", 1.0),
          ("dram write GB", "dram__bytes_write.sum", 1.0),
          ("dram % peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
          ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
          ("regs", "launch__registers_per_thread", 1.0),
          ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
          ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
          ("IPC", "sm__inst_executed.avg.per_cycle_elapsed", 1.0),
          ("threads/inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
          ("warp inst (G)", "smsp__inst_executed.sum", 1e-9)]


def full_table(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["| kernel | " + " | ".join(f[0] for f in FIELDS) + " |", "|---|" + "---|" * len(FIELDS)]
    for r in rows[2:]:
        cells = []
        for label, metric, scale in FIELDS:
            v = r[idx[metric]] if metric in idx else ""
            try:
                x = float(v.replace(",", "")) * scale
                unit = units[idx[metric]]
                if label.endswith("GB") and unit == "Mbyte":
                    x /= 1e3
                if label == "time ms" and unit == "us":
                    x /= 1e3
                cells.append("%.3f" % x if x < 100 else "%.1f" % x)
            except ValueError:
                cells.append(v)
        out.append("| %s | %s |" % (short(r[idx["Kernel Name"]]), " | ".join(cells)))
    return "\n".join(out)


if __name__ == "__main__":
    table, total = launch_table(sys.argv[1])
    print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), share of GPU time\n")
    print(table)
    print("\nTotal GPU time in the list: %.1f ms\n" % (total / 1e3))
    if len(sys.argv) > 2:
        print("## Full capture (`ncu --set full --clock-control none --import-source on`), per launch\n")
        print(full_table(sys.argv[2]))
