#!/usr/bin/env python
"""Per-source-line instruction counts of one profiled launch: joins the SASS page of an ncu report (complete executed
counts per instruction, no source lines) with `nvdisasm --print-line-info` of the same cubin (source line per instruction).

    nvdisasm --print-line-info X.cubin > all.sass      (cut to the function's section)
    sass_by_line.py func.sass src.csv LAUNCH UNITS [TOP]
"""
import collections
import csv
import re
import sys


def main():
    sass, src, launch, units = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    ins, cur = [], ("", 0)
    for l in open(sass):
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((cur, m.group(2).strip()))
    rows = list(csv.reader(open(src)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    a, b = secs[2 * launch], secs[2 * launch + 1]
    hdr = rows[a + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    if len(ex) != len(ins):
        raise SystemExit("instruction counts differ: {} in the report, {} in the disassembly".format(len(ex), len(ins)))
    agg = collections.defaultdict(collections.Counter)
    for (key, text), r in zip(ins, ex):
        op = text.split()[1] if text.startswith("@") else text.split()[0]
        if op.split(".")[0] != r[ix["Source"]].split()[1 if r[ix["Source"]].strip().startswith("@") else 0].split(".")[0]:
            raise SystemExit("the disassembly is not the profiled code: {} vs {}".format(text, r[ix["Source"]]))
        n = int(r[ix["Instructions Executed"]])
        c = agg[key]
        c["ins"] += n
        c["smp"] += int(r[ix["# Samples"]])
        base = op.split(".")[0]
        if base in ("DADD", "DFMA", "DMUL", "DSETP"):
            c["fp64"] += n
        elif base in ("IMAD", "MOV") and (".MOV" in op or base == "MOV"):
            c["mov"] += n
        elif base in ("SHFL",):
            c["shfl"] += n
        elif base in ("LDS", "STS", "LDSM"):
            c["smem"] += n
        elif base in ("LDC", "LDCU", "ULDC"):
            c["ldc"] += n
        else:
            c["other"] += n
    tot = collections.Counter()
    for c in agg.values():
        tot.update(c)
    print("per unit: " + " ".join("{} {:.1f}".format(k, tot[k] / units) for k in ("ins", "fp64", "mov", "shfl", "smem", "ldc", "other")))
    for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["ins"])[:top]:
        print("{:>22s}:{:<4d} ins {:6.1f} fp64 {:6.1f} mov {:5.1f} shfl {:5.1f} smem {:5.1f} ldc {:5.1f} other {:6.1f}  smp {:4.1f}%".format(
            key[0][-22:], key[1], c["ins"] / units, c["fp64"] / units, c["mov"] / units, c["shfl"] / units,
            c["smem"] / units, c["ldc"] / units, c["other"] / units, 100.0 * c["smp"] / max(1, tot["smp"])))


if __name__ == "__main__":
    main()
