#!/bin/bash
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r2g}
for spec in "busiest 37" "low 9"; do
  set -- $spec
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s $2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_probe.py --steps $(($2 + 2)) --variants evl > gpurun_out/lines_$1_$TAG.log 2>&1
  python tools/ncu_sass_dump.py /tmp/prof_$1.ncu-rep gpurun_out/${TAG}_sass_$1.csv.gz
done
du -sh gpurun_out
