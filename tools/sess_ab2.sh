#!/bin/bash
# In-session A/B of builds: LIBS="name1 name2" (ev2gym_b200/csrc/libev2b_<name>.so; "cur" = libev2b.so), two rounds each.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-ab2}
for round in ${ROUNDS:-1 2}; do
  for name in ${LIBS:-prev cur}; do
    if [ "$name" = cur ]; then unset EV2B_LIB; else export EV2B_LIB=$PWD/ev2gym_b200/csrc/libev2b_$name.so; fi
    timeout 300 python tools/ab_kernels.py --workloads ${WL:-c3,c4,c5} --variants ${VARIANTS:-evl} --out gpurun_out/ab_${TAG}_${name}_$round.json > gpurun_out/ab_${TAG}_${name}_$round.log 2>&1
    python - <<PY
import json
for l in open("gpurun_out/ab_${TAG}_${name}_$round.json"):
    d = json.loads(l); print("$name", $round, d["workload"], d["us_per_launch_episode"], "idle", d["idle_us"], "busy", d["busy_us"], "busiest", d["busiest_us"])
PY
  done
done
