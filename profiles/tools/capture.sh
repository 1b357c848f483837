#!/bin/bash
# ncu captures of the kernels the roofline entries of bench.py refer to (run on the GPU box, after bench.py itself has
# exited 0 without ncu).  usage: profiles/tools/capture.sh TAG   -> gpurun_out/TAG_*_raw.csv / _src.csv (+ one .ncu-rep)
# gpurun brings back at most 64 MiB: the reports are turned into their CSV pages here and only k_cycle1's is kept.
set -u
TAG=${1:-cap}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-plain"
pages() {   # report -> raw + source CSV pages; $2 = keep the report
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_src.csv 2>/dev/null
  [ "${2:-}" = keep ] || rm -f $OUT/$1.ncu-rep
}
# k_cycle1: first and second noisy cycle of the timed step (launches 4 and 5; 3 is its preparation cycle)
$NCU -k regex:k_cycle1 --launch-skip 4 --launch-count 2 -o $OUT/${TAG}_cycle1 $B --no-dense > $OUT/${TAG}_cycle1.log 2>&1
pages ${TAG}_cycle1 keep
# k_plan: first noisy cycle of the timed step
$NCU -k regex:k_plan --launch-skip 4 --launch-count 1 -o $OUT/${TAG}_plan $B --no-dense > $OUT/${TAG}_plan.log 2>&1
pages ${TAG}_plan
# full-state mode (S17_DENSE=1 inside bench.py's dense section): sweeps 5-7 of the first noisy cycle (the last one feeds the Measure), two collapses, one sampling
D="$B --dense-shots 4096 --dense-reps 1"
$NCU -k regex:k_layer2 --launch-skip 37 --launch-count 3 -o $OUT/${TAG}_layer2 $D > $OUT/${TAG}_layer2.log 2>&1
pages ${TAG}_layer2
$NCU -k regex:k_collapse --launch-skip 2 --launch-count 2 -o $OUT/${TAG}_collapse $D > $OUT/${TAG}_collapse.log 2>&1
pages ${TAG}_collapse
$NCU -k regex:k_measure --launch-skip 3 --launch-count 1 -o $OUT/${TAG}_measure $D > $OUT/${TAG}_measure.log 2>&1
pages ${TAG}_measure
du -sh $OUT; ls -la $OUT/${TAG}_*
