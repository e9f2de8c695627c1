#!/bin/bash
# usage: scripts/ncu_one.sh <tag> <H> <W> <engine> <noise> [fdt]   -> gpurun_out/<tag>.ncu-rep + summary
set -e
tag=$1; shift
python scripts/quick_bench.py "$@" 12 > gpurun_out/${tag}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 8 -c 1 -f -o gpurun_out/${tag} \
    python scripts/quick_bench.py "$@" 12 > gpurun_out/${tag}_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}.ncu-rep 40 > gpurun_out/${tag}_summary.txt 2>&1
