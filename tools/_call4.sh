mkdir -p gpurun_out
T="tests/test_gpu_two_phase.py tests/test_gpu_slabs.py"
echo "== stash, x32+x4"; LBM_TP_STASH=1 timeout 300 python -m pytest $T -m gpu -q -x 2>&1 | tail -4
echo "== stash, x4 pieces"; LBM_TEST_LIB=lattice-boltzmann-method_b200/liblbm_b200_x4.so LBM_TP_STASH=1 timeout 300 python -m pytest $T -m gpu -q -x 2>&1 | tail -4
echo "== stash NS=2"; LBM_TP_STASH=1 LBM_TP_NS=2 timeout 300 python -m pytest tests/test_gpu_two_phase.py -m gpu -q -x 2>&1 | tail -2
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-44s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
for lib in liblbm_b200.so liblbm_b200_x4.so; do
for cfg in "LBM_TP_STASH=1 LBM_TP_NS=2" "LBM_TP_STASH=1 LBM_TP_NS=3" "LBM_TP_STASH=1 LBM_TP_NS=4"; do
  n=$(echo "$lib $cfg" | tr ' =.' '___')
  env $cfg timeout 200 python bench.py --lib lattice-boltzmann-method_b200/$lib --workload mrtcg_rt_weak $B 2>gpurun_out/c4_err.txt | tail -1 > gpurun_out/c4_mrtcg_$n.json; show "mrtcg 8192x16384 $lib $cfg" gpurun_out/c4_mrtcg_$n.json
  env $cfg timeout 200 python bench.py --lib lattice-boltzmann-method_b200/$lib --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c4_err.txt | tail -1 > gpurun_out/c4_rk_$n.json; show "rk 4096^2 $lib $cfg" gpurun_out/c4_rk_$n.json
done; done
tail -3 gpurun_out/c4_err.txt
echo "== bench-scale parity (default path)"; ( time timeout 900 python -m pytest tests/test_gpu_bench_scale.py -m gpu -q --durations=8 2>&1 | tail -16 ) 2>&1 | tail -22
CMD="python bench.py --workload mrtcg_rt_weak --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
LBM_TP_STASH=1 $CMD > gpurun_out/c4_plain.log 2>&1 && LBM_TP_STASH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tp_staged -s 2 -c 1 -o gpurun_out/r02_ncu_mrtcg_stash -f $CMD > gpurun_out/c4_ncu.log 2>&1
tail -2 gpurun_out/c4_ncu.log | cut -c1-200
