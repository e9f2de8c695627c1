#!/bin/bash
# development aid: bench.py value / kernel time for library variants, interleaved reps
# usage: ab_bench.sh out_file steps variant...   ("main" = in-tree)
out=$1; steps=$2; shift; shift
for rep in 1 2 3 4; do
  for v in "$@"; do
    if [ "$v" == "main" ]; then unset V2E_LIB; else export V2E_LIB=$PWD/variants/lib_$v.so; fi
    python bench.py --steps $steps --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', d['value'], d['ms_per_step'], r['frac'], r['kernel_ms_per_pack'])" >> $out
  done
done
