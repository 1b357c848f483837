#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass) per source line.

usage: ncu_by_line.py file.csv [n_units] [top]
Prints, per file:line, warp instructions / unit, FP64 instructions, shared wavefronts and stall samples."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    n = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(path)))
    fname = ""
    hdr = None
    agg = collections.defaultdict(lambda: collections.Counter())
    text = {}
    cur = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ix = {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0] == "" or not r[0].isdigit():
            continue                                   # SASS sub-rows (partly elided): the line row carries the totals
        cur = (fname, int(r[0]))
        text[cur] = r[1].strip()

        def val(name):
            v = r[ix[name]]
            return int(v) if v.lstrip("-").isdigit() else 0
        a = agg[cur]
        a["ins"] += val("Instructions Executed")
        a["wave"] += val("L1 Wavefronts Shared")
        a["xwave"] += val("L1 Wavefronts Shared Excessive")
        a["samples"] += val("# Samples")
    tot = collections.Counter()
    for a in agg.values():
        tot.update(a)
    print("total per unit: ins {:.1f} fp64 {:.1f} shfl {:.1f} wave {:.1f} xwave {:.1f} samples {}".format(
        tot["ins"] / n, tot["fp64"] / n, tot["shfl"] / n, tot["wave"] / n, tot["xwave"] / n, tot["samples"]))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        print("{:>16s}:{:<4d} ins {:7.1f} fp64 {:6.1f} shfl {:5.1f} wave {:6.1f} x {:5.1f} smp {:5.1f}%  {}".format(
            key[0][-16:], key[1], a["ins"] / n, a["fp64"] / n, a["shfl"] / n, a["wave"] / n, a["xwave"] / n,
            100.0 * a["samples"] / max(1, tot["samples"]), text.get(key, "")[:70]))


if __name__ == "__main__":
    main()
