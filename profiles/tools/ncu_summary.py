#!/usr/bin/env python
"""One line per profiled launch of an ncu raw page (ncu -i X.ncu-rep --page raw --csv): the metrics the roofline entries use."""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conflicts"),
        ("smsp__inst_executed.sum", "warp_instr"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, unit = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print("| # | kernel | " + " | ".join(k for _, k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for n, r in enumerate(rows[2:]):
    name = r[ix["Kernel Name"]].split("(")[0].replace("s17::", "").replace("void ", "")
    cells = []
    for key, _ in KEYS:
        if key in ix:
            v, u = r[ix[key]], unit[ix[key]]
            try:
                v = "{:.4g}".format(float(v.replace(",", "")))
            except ValueError:
                pass
            cells.append((v + " " + u).strip())
        else:
            cells.append("-")
    print("| {} | {} | ".format(n, name) + " | ".join(cells) + " |")
