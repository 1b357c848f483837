#!/bin/bash
# A/B of libs17.so builds on one box: profiles/tools/ab.sh TAG lib1.so lib2.so ...  -> gpurun_out/TAG_ab.jsonl
# (one short un-memoised bench per library through S17_LIB; same seeds, so the counters must agree)
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_ab.jsonl
for lib in "$@"; do
  S17_LIB=$PWD/$lib python bench.py --no-cpu --no-dense --no-plain --steps 10 --warmup 3 2> $OUT/${TAG}_ab_err.log | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({'lib': '$lib', 'value': d['value'], 'e2e': d['e2e']['value'], 'ms_per_step': d['ms_per_step'], 'counts': d['counts'], 'gate_share': d['roofline'].get('gate_kernel_share_of_step'), 'clocks': d.get('clocks')}))" >> $OUT/${TAG}_ab.jsonl
done
cat $OUT/${TAG}_ab.jsonl
