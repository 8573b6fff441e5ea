mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
L=gpurun_out/steps_r2c.log; rm -f $L
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method=thread > gpurun_out/test_gpu_r2c.log 2>&1; echo "gpu tests rc=$?" >> $L
timeout 200 python tools/ab_kernels.py --workloads c3,c4 --variants evl:G=1,evl:G=2,evl:G=4 --out gpurun_out/ab_r2c.json > gpurun_out/ab_r2c.log 2>&1; echo "ab rc=$?" >> $L
for mb in 7 6; do EV2B_LIB=$PWD/ev2gym_b200/csrc/libev2b_mb$mb.so timeout 150 python tools/ab_kernels.py --workloads c3,c4 --variants evl:G=1,evl:G=2 --out gpurun_out/ab_r2c_mb$mb.json > gpurun_out/ab_r2c_mb$mb.log 2>&1; echo "ab mb$mb rc=$?" >> $L; done
timeout 150 python tools/ab_kernels.py --workloads c5,c3-1k --variants percharger,evl:G=1,evl:G=2 --out gpurun_out/ab_r2c_c5.json > gpurun_out/ab_r2c_c5.log 2>&1; echo "ab c5 rc=$?" >> $L
timeout 300 python bench.py > gpurun_out/bench_c3_r2c.json 2> gpurun_out/bench_c3_r2c.err; echo "bench rc=$?" >> $L
for G in 1 2; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o /tmp/prof_busy_g$G python tools/ncu_probe.py --steps 30 --variants evl:G=$G > gpurun_out/prof_busy_r2c_g$G.log 2>&1
python tools/ncu_summary.py /tmp/prof_busy_g$G.ncu-rep > gpurun_out/r2c_evl_g${G}_ncu_busy_step.txt 2>&1
python tools/ncu_lines.py /tmp/prof_busy_g$G.ncu-rep ev2gym_b200/csrc/libev2b.so evl_step_kernelIfLi2ELb1ELi${G}ELb0 60 > gpurun_out/r2c_evl_g${G}_ncu_busy_lines.txt 2>&1
echo "prof G=$G rc=$?" >> $L
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 3 -c 1 -o /tmp/prof_idle_g1 python tools/ncu_probe.py --steps 6 --variants evl:G=1 > gpurun_out/prof_idle_r2c_g1.log 2>&1
python tools/ncu_summary.py /tmp/prof_idle_g1.ncu-rep > gpurun_out/r2c_evl_g1_ncu_idle_step.txt 2>&1
python tools/ncu_lines.py /tmp/prof_idle_g1.ncu-rep ev2gym_b200/csrc/libev2b.so evl_step_kernelIfLi2ELb1ELi1ELb0 40 > gpurun_out/r2c_evl_g1_ncu_idle_lines.txt 2>&1
cp /tmp/prof_busy_g1.ncu-rep gpurun_out/prof_busy_r2c_g1.ncu-rep
cat $L; tail -5 gpurun_out/test_gpu_r2c.log; du -sh gpurun_out
