#!/bin/bash
# development aid: fixed-point engine, library-variant sweep on the GPU box (variants prebuilt
# with scripts/build_variant.sh; "default" = the in-tree library)
for v in "$@"; do
  if [ "$v" = default ]; then unset V2E_LIB; else export V2E_LIB=$PWD/v2e2v_b200/_build/lib_$v.so; fi
  echo "=== variant $v"
  for n in off philox; do
    timeout 120 python scripts/quick_bench.py 720 1280 fixpoint $n f32 40 2>&1 | grep -v "host enqueue"
    timeout 120 python scripts/timeline_fx.py 720 1280 $n 2>&1 | tail -${FX_TL_LINES:-11}
  done
done
