mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_r2b.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/steps_r2b.log
timeout 300 python tools/ab_kernels.py --workloads c3,c4,c3-1k,c2 --variants percharger,evl:G=1,evl:G=2,evl:G=4 --out gpurun_out/ab_r2b.json > gpurun_out/ab_r2b.log 2>&1; echo "ab rc=$?" >> gpurun_out/steps_r2b.log
for mb in 7 6; do EV2B_LIB=$PWD/ev2gym_b200/csrc/libev2b_mb$mb.so timeout 200 python tools/ab_kernels.py --workloads c3,c4 --variants evl:G=1,evl:G=2 --out gpurun_out/ab_r2b_mb$mb.json > gpurun_out/ab_r2b_mb$mb.log 2>&1; echo "ab mb$mb rc=$?" >> gpurun_out/steps_r2b.log; done
timeout 300 python bench.py > gpurun_out/bench_c3_r2b.json 2> gpurun_out/bench_c3_r2b.err; echo "bench rc=$?" >> gpurun_out/steps_r2b.log
for G in 1 2; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 3 -c 1 -o gpurun_out/prof_idle_r2b_g$G python tools/ncu_probe.py --steps 6 --variants evl:G=$G > gpurun_out/prof_idle_r2b_g$G.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o gpurun_out/prof_busy_r2b_g$G python tools/ncu_probe.py --steps 30 --variants evl:G=$G > gpurun_out/prof_busy_r2b_g$G.log 2>&1
echo "prof G=$G rc=$?" >> gpurun_out/steps_r2b.log
done
cat gpurun_out/steps_r2b.log; tail -5 gpurun_out/test_gpu_r2b.log
