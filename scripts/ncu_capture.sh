#!/bin/bash
# usage: scripts/ncu_capture.sh <tag> <kernel regex> [env assignments...]   (run under gpurun; writes gpurun_out/<tag>.ncu-rep + CSV pages)
tag=$1; kre=$2; shift 2
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on -k regex:$kre -s 2 -c 1 -f -o gpurun_out/$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/$tag.src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/$tag.raw.csv gpurun_out/$tag.src.csv > gpurun_out/$tag.summary.txt 2>&1
python scripts/ncu_compact.py gpurun_out/$tag.src.csv gpurun_out/$tag.pc.csv.gz
rm -f gpurun_out/$tag.ncu-rep gpurun_out/$tag.src.csv     # too big to travel back (64 MiB limit); the compact per-PC table stays
tail -2 gpurun_out/$tag.log
