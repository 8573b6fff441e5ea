#!/bin/bash
# ncu --set full of evl_step_kernel at chosen episode steps of c3: summary, per-source-line and per-SASS tables (text only).
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r2g}
for spec in ${SPECS:-busiest:37 low:9}; do
  name=${spec%%:*}; step=${spec##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s $step -c 1 -f -o /tmp/prof_$name python tools/ncu_probe.py --steps $(($step + 2)) --variants evl > gpurun_out/lines_${name}_$TAG.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$name.ncu-rep > gpurun_out/${TAG}_evl_ncu_${name}_step.txt 2>&1
  python tools/ncu_lines.py /tmp/prof_$name.ncu-rep ev2gym_b200/csrc/libev2b.so evl_step_kernelIfLi2ELb1ELi1ELb0ELb0ELi128 120 > gpurun_out/${TAG}_evl_lines_$name.txt 2>&1
  python tools/ncu_sass_dump.py /tmp/prof_$name.ncu-rep gpurun_out/${TAG}_sass_$name.csv.gz ${DUMP_ALL:+--all}
done
du -sh gpurun_out
