set -x
B="python bench.py --steps 2 --warmup 3 --no-extra-configs --no-cpu-baseline --no-gpu-baseline"
$B > gpurun_out/r2f_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2f_launches.csv $B > gpurun_out/r2f_ncu_bench.log 2>&1
python tools/bench_one.py shared > gpurun_out/r2f_plain_one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rr2t -c 1 -o gpurun_out/r2f_prof_rr2t_residual python tools/bench_one.py shared > gpurun_out/r2f_ncu_one.log 2>&1
python tools/profile_torch.py > gpurun_out/r2f_prof_step.txt 2>&1
tail -2 gpurun_out/r2f_ncu_bench.log gpurun_out/r2f_ncu_one.log
head -12 gpurun_out/r2f_prof_step.txt
