#!/bin/bash
# Last check of a round: the GPU test suite, smoke(), the default bench line (c3, all extras) and the 1024-env workload.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-final}
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/test_gpu_$TAG.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/test_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 500 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench rc=$?"
timeout 200 python bench.py --workload c2 --no-extras --min-seconds 0.4 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
du -sh gpurun_out
