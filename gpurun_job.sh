set -x
# launch lists (serialised, cold-cache per-launch times): shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_mrtcg.csv python bench.py --workload mrtcg_rt --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l_mrt.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_rk.csv python bench.py --workload rk_droplet --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l_rk.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sed.csv python bench.py --workload sedimentation --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l_sed.log 2>&1
# full captures of the dominant kernels at the bench sizes
ncu --set full --clock-control none --import-source on -k regex:k_tp_fused -s 3 -c 1 -o gpurun_out/prof_mrtcg_16384 python bench.py --workload mrtcg_rt --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f_mrt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tp_fused -s 3 -c 1 -o gpurun_out/prof_rk_4096 python bench.py --workload rk_droplet --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f_rk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bgk_interior -s 8 -c 1 -o gpurun_out/prof_sed python bench.py --workload sedimentation --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f_sed.log 2>&1
ls -la gpurun_out
