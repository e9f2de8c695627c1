#!/bin/bash
# development aid: does the clock sampler disturb the timed region?  value of the default bench, 3 runs each
for rep in 1 2 3; do
  for mode in none nvml2 nvml10 nvml50 smi; do
    unset BENCH_NO_SAMPLER BENCH_SAMPLER_MS BENCH_SAMPLER_SMI
    case $mode in
      none) export BENCH_NO_SAMPLER=1;;
      nvml2) export BENCH_SAMPLER_MS=2;;
      nvml10) export BENCH_SAMPLER_MS=10;;
      nvml50) export BENCH_SAMPLER_MS=50;;
      smi) export BENCH_SAMPLER_SMI=1;;
    esac
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$mode', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_pack'], d['clocks'])"
  done
done
