#!/usr/bin/env python
"""Per-unit instruction counts and DRAM traffic of one profiled launch, from the two CSV pages of an ncu report:

    ncu -i X.ncu-rep --page raw --csv > raw.csv
    ncu -i X.ncu-rep --page source --csv > src.csv
    ncu_counts.py raw.csv src.csv --launch 1 --units 524288 --source "profiles/....ncu-rep (what was profiled)" > counts.json

units = what one launch processed (k_cycle1: shot-cycles = active shots of that launch).  bench.py reads the JSON
(profiles/k_cycle1_counts.json) for the FP64-pipe floor and the measured DRAM traffic of its roofline entry."""
import argparse
import collections
import csv
import json

FP64 = ("DADD", "DFMA", "DMUL", "DSETP", "DMNMX")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("src")
    ap.add_argument("--launch", type=int, default=0, help="index of the launch in the report")
    ap.add_argument("--units", type=float, required=True)
    ap.add_argument("--source", default="")
    a = ap.parse_args()

    rows = list(csv.reader(open(a.raw)))
    hdr, unit = rows[0], rows[1]
    r = rows[2 + a.launch]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(name):
        v = float(r[ix[name]].replace(",", ""))
        u = unit[ix[name]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
                 "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        return v * scale

    srows = list(csv.reader(open(a.src)))
    secs = [i for i, x in enumerate(srows) if x and x[0] == "Kernel Name"] + [len(srows)]
    want = int(round(val("smsp__inst_executed.sum")))
    ops = None
    for s0, s1 in zip(secs, secs[1:]):            # the page can list a launch more than once: match by instruction total
        shdr = srows[s0 + 1]
        six = {h: i for i, h in enumerate(shdr)}
        cand = collections.Counter()
        waves = xwaves = 0
        for x in srows[s0 + 2:s1]:
            if len(x) != len(shdr):
                continue
            sass = x[six["Source"]].split()
            if not sass:
                continue
            op = (sass[1] if sass[0].startswith("@") else sass[0]).split(".")[0]
            cand[op] += int(x[six["Instructions Executed"]])
            waves += int(x[six["L1 Wavefronts Shared"]])
            xwaves += int(x[six["L1 Wavefronts Shared Excessive"]])
        if abs(sum(cand.values()) - want) <= 1e-5 * want:      # (the two pages come from different replay passes)
            ops = cand
            break
    if ops is None:
        raise SystemExit("no section of the source page matches launch {} ({} warp instructions)".format(a.launch, want))
    n = a.units
    out = {
        "kernel": r[ix["Kernel Name"]].split("(")[0],
        "source": a.source,
        "units_in_launch": n,
        "duration_ms": 1e3 * val("gpu__time_duration.sum"),
        "warp_instr_per_shot_cycle": val("smsp__inst_executed.sum") / n,
        "fp64_warp_instr_per_shot_cycle": sum(ops[o] for o in FP64) / n,
        "opcodes_per_shot_cycle": {k: round(v / n, 2) for k, v in ops.most_common(16)},
        "shared_wavefronts_per_shot_cycle": waves / n,
        "shared_wavefronts_excess_per_shot_cycle": xwaves / n,
        "dram_bytes_per_shot_cycle": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / n,
        "fp64_pipe_active_pct": float(r[ix["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
        "issue_active_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
        "shared_wavefronts_pct": float(r[ix["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]]),
        "registers_per_thread": int(float(r[ix["launch__registers_per_thread"]])),
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
