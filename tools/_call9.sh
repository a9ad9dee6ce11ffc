mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-50s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
i=0
for rep in 1 2; do
for lib in liblbm_b200_old.so liblbm_b200.so liblbm_b200_modwrap.so liblbm_b200_nocarve.so; do i=$((i+1))
  timeout 200 python bench.py --lib lattice-boltzmann-method_b200/$lib --workload mrtcg_rt_weak $B 2>gpurun_out/c9_err.txt | tail -1 > gpurun_out/c9_m_$i.json; show "mrtcg stash $lib" gpurun_out/c9_m_$i.json
done; done
for lib in liblbm_b200_old.so liblbm_b200.so; do i=$((i+1))
  LBM_TP_RPB=128 timeout 200 python bench.py --lib lattice-boltzmann-method_b200/$lib --workload mrtcg_rt_weak $B 2>gpurun_out/c9_err.txt | tail -1 > gpurun_out/c9_m_$i.json; show "mrtcg stash RPB=128 $lib" gpurun_out/c9_m_$i.json
  LBM_TP_STASH=0 timeout 200 python bench.py --lib lattice-boltzmann-method_b200/$lib --workload mrtcg_rt_weak $B 2>gpurun_out/c9_err.txt | tail -1 > gpurun_out/c9_n_$i.json; show "mrtcg STASH=0 $lib" gpurun_out/c9_n_$i.json
  timeout 200 python bench.py --lib lattice-boltzmann-method_b200/$lib --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c9_err.txt | tail -1 > gpurun_out/c9_rk_$i.json; show "rk default $lib" gpurun_out/c9_rk_$i.json
done
tail -2 gpurun_out/c9_err.txt
