#!/bin/bash
# usage: tl_variants.sh out_prefix variant...   (timeline + clip probe for each library variant; "main" = in-tree)
pre=$1; shift
for v in "$@"; do
  if [ "$v" == "main" ]; then unset V2E_LIB; else export V2E_LIB=$PWD/variants/lib_$v.so; fi
  python scripts/timeline_fx.py 720 1280 philox > gpurun_out/${pre}_tl_$v.txt 2>&1
  python scripts/clip_probe.py philox fixpoint 50 > gpurun_out/${pre}_clip_$v.txt 2>&1
done
