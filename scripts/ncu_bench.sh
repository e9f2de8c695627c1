#!/bin/bash
# usage: scripts/ncu_bench.sh <tag> [bench args]  -> gpurun_out/<tag>_{bench.json,launches.csv,chain.ncu-rep,chain_summary.txt}
# (1) bench.py plain, (2) the ncu launch list of the same command, (3) one --set full capture of a
# chained launch of the specialised kernel (the 7th launch: the first chain of 16 packs of the warm-up clip, whose chunks are 1, 16, 16, ... packs).
tag=$1; shift
python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py "$@" > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fixpoint_spec -s 6 -c 1 -f -o gpurun_out/${tag}_chain \
    python bench.py "$@" > gpurun_out/${tag}_ncu_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_chain.ncu-rep 40 > gpurun_out/${tag}_chain_summary.txt 2>&1
python scripts/ncu_regions.py gpurun_out/${tag}_chain.ncu-rep > gpurun_out/${tag}_chain_regions.txt 2>&1
python scripts/ncu_lines.py gpurun_out/${tag}_chain.ncu-rep > gpurun_out/${tag}_chain_lines.txt 2>&1
