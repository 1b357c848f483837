#!/usr/bin/env python
"""k_cycle1's executed warp instructions per shot-cycle by REGION of the kernel (per-shot prologue, transform, event scan,
table pointers, the half cycle itself, store ...) and, inside a region, by the innermost source line.

Joins the SASS page of an ncu report (executed count and stall samples per instruction) with
`nvdisasm --print-line-info-inline` of the same build (the inlining chain of every instruction: the OUTERMOST line inside
the kernel body says which region an inlined helper was called from).

    cuobjdump -xelf all surface_17_iqt_b200/libs17.so && nvdisasm --print-line-info-inline *.cubin > all_inl.sass
    ncu_regions.py all_inl.sass SRC.csv UNITS [LAUNCH] [REGION]
"""
import collections
import csv
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "..", "surface_17_iqt_b200", "csrc", "s17_diag.cuh")
FUNC = ".text._ZN3s178k_cycle1ILb0"


def main():
    sass, src, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    want = sys.argv[5] if len(sys.argv) > 5 else None
    lines = open(sass).read().split("\n")
    start = [i for i, l in enumerate(lines) if l.startswith(FUNC)][0]
    end = [i for i, l in enumerate(lines) if l.startswith("//-") and ".text." in l and i > start][0]
    ins, pend, cur = [], [], []
    for l in lines[start:end]:
        m = re.search(r'//## File "[^"]*/([^"/]+)", line (\d+)( inlined at "[^"]*/([^"/]+)", line (\d+))?', l)
        if m:
            pend.append(m)
            continue
        m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m2:
            if pend:
                cur = [(q.group(1), int(q.group(2))) for q in pend]
                if pend[-1].group(3):
                    cur.append((pend[-1].group(4), int(pend[-1].group(5))))
                pend = []
            ins.append((cur, m2.group(2).strip()))
    rows = list(csv.reader(open(src)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    a, b = secs[2 * launch], secs[2 * launch + 1]
    hdr = rows[a + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    if len(ex) != len(ins):
        raise SystemExit("instruction counts differ: {} in the report, {} in the disassembly".format(len(ex), len(ins)))
    srcl = open(SRC).read().split("\n")

    def find(s, first=0):
        for i in range(first, len(srcl)):
            if s in srcl[i]:
                return i + 1
        raise KeyError(s)

    kstart = find("k_cycle1(const __grid_constant__")
    half = find("dg_half1(p, tk0, tk1")
    marks = [("kernel prologue + prefetch lambda", kstart),
             ("per shot: list, uniforms, load", find("for (uint32_t item = first")),
             ("loop over the halves", find("for (int h = 0; h < 2; ++h) {", find("for (uint32_t item = first"))),
             ("transform", find("// Walsh-Hadamard transform of the data register: the register bits")),
             ("event scan", find("// This shot's events in this half cycle")),
             ("private entries", find("if (dirty4) {")),
             ("table pointers, signs", find("const unsigned char *clean_h")),
             ("half cycle (dg_half1)", half),
             ("after the half", find("if (h == 0) carry = fw", half)),
             ("read-out draw", find("if (ma.sample_data) {")),
             ("stage + store", find("// stage the (already renormalised) block")),
             ("kernel epilogue", find("if (lane == 0) {", find("entry1 = entry2;")))]

    def outer(chain):
        for f, l in reversed(chain):
            if f == "s17_diag.cuh" and l >= kstart:
                return l
        return -1

    def region(l):
        r = "?"
        for name, ln in marks:
            if l >= ln:
                r = name
        return r

    agg, sub = collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
    for (chain, text), r in zip(ins, ex):
        n, smp = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        op = text.split()[1] if text.startswith("@") else text.split()[0]
        base = op.split(".")[0]
        cat = ("fp64" if base in ("DADD", "DFMA", "DMUL", "DSETP") else "mov" if (base == "MOV" or ".MOV" in op) else
               "shfl" if base == "SHFL" else "smem" if base in ("LDS", "STS") else "other")
        reg = region(outer(chain))
        inner = [l for f, l in chain if f == "s17_diag.cuh"]
        for c in (agg[reg], sub[(reg, inner[0] if inner else -1)]):
            c["ins"] += n
            c[cat] += n
            c["smp"] += smp
    ts = sum(c["smp"] for c in agg.values())
    fmt = "ins {:7.1f} fp64 {:6.1f} mov {:5.1f} shfl {:5.1f} smem {:5.1f} other {:6.1f}  samples {:4.1f}%"

    def show(c):
        return fmt.format(c["ins"] / units, c["fp64"] / units, c["mov"] / units, c["shfl"] / units, c["smem"] / units,
                          c["other"] / units, 100.0 * c["smp"] / max(1, ts))

    print("warp instructions per shot-cycle: {:.1f}".format(sum(c["ins"] for c in agg.values()) / units))
    for name in [m[0] for m in marks] + ["?"]:
        if agg[name]["ins"]:
            print("{:<34s} {}".format(name, show(agg[name])))
    if want:
        print("-- " + want)
        for k in sorted(sub):
            if k[0] == want and sub[k]["ins"] / units >= 4:
                print("{:>5d} {}  {}".format(k[1], show(sub[k]), srcl[k[1] - 1].strip()[:60] if k[1] > 0 else ""))


if __name__ == "__main__":
    main()
