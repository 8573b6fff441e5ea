#!/bin/bash
# DRAM bytes per launch of the dominant kernel at the busiest c3 step, stamped with the hash of the CUDA sources
# (-> profiles/roofline_traffic.json, which bench.py reads for roofline.traffic / traffic_is_this_build).
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-final}
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 37 -c 1 -f -o /tmp/prof_busy python tools/ncu_probe.py --steps 39 --variants evl > gpurun_out/prof_busy_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/prof_busy.ncu-rep > gpurun_out/${TAG}_evl_ncu_busiest_step.txt 2>&1; echo "prof busy rc=$?"
python - <<PY > gpurun_out/roofline_traffic_$TAG.json
import json, re, sys
sys.path.insert(0, ".")
from bench import source_sha
txt = open("gpurun_out/${TAG}_evl_ncu_busiest_step.txt").read()
def val(name):
    m = re.search(name + r"\s+([0-9.]+) (\w+)", txt)
    return float(m.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
print(json.dumps({"c3": {"dram_bytes_per_launch": b, "source_sha": source_sha(),
      "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of ev2b::evl_step_kernel<float,2,true,1,false,false> at episode step 37 of 112 (the busiest: 130 of 200 ports occupied), c3, profiles/${TAG}_evl_ncu_busiest_step.txt"}}, indent=1))
PY
cat gpurun_out/roofline_traffic_$TAG.json
