# final round-2 evidence (gpurun -- bash tools/prof_final.sh): each ncu pass runs only after the
# same command exited 0 without ncu
set -x
B="python bench.py --steps 2 --warmup 3 --no-extra-configs --no-cpu-baseline --no-gpu-baseline"
$B > gpurun_out/r2g_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2g_launches.csv $B > gpurun_out/r2g_ncu_bench.log 2>&1
python tools/bench_one.py > gpurun_out/r2g_bench_one.txt 2>&1; cat gpurun_out/r2g_bench_one.txt
python tools/bench_one.py dgrad_gate > gpurun_out/r2g_plain_one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"rr2t|reflect_border" -c 2 -o gpurun_out/r2g_prof_dgrad_gate python tools/bench_one.py dgrad_gate > gpurun_out/r2g_ncu_one.log 2>&1
python tools/bench_one.py shared > gpurun_out/r2g_plain_one2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rr2t -c 1 -o gpurun_out/r2g_prof_rr2t_residual python tools/bench_one.py shared > gpurun_out/r2g_ncu_one2.log 2>&1
python tools/bench_border.py > gpurun_out/r2g_bench_border.txt 2>&1; cat gpurun_out/r2g_bench_border.txt
python tools/profile_torch.py > gpurun_out/r2g_prof_step.txt 2>&1
python tools/profile_torch.py --size 256 > gpurun_out/r2g_prof_step_256.txt 2>&1
tail -2 gpurun_out/r2g_ncu_bench.log gpurun_out/r2g_ncu_one.log gpurun_out/r2g_ncu_one2.log
ls -la gpurun_out/*.ncu-rep
