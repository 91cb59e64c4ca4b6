#!/bin/bash
# ncu --set full of selected kernels of one bench step.  bash profiles/_ncu.sh <tag> <kernel-regex> <skip> <count>
tag=$1; rx=$2; skip=$3; cnt=$4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o /tmp/${tag}_full \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-strong > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > /tmp/${tag}_full_raw.csv && python profiles/ncu_select.py /tmp/${tag}_full_raw.csv gpurun_out/${tag}_full_sel.csv
ls -la /tmp/${tag}_full.ncu-rep
