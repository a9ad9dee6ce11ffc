mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_two_phase.py tests/test_gpu_slabs.py tests/test_gpu_long_horizon.py tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -5
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-34s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
for cfg in "LBM_TP_STAGED=0" "LBM_TP_NS=4" "LBM_TP_NS=5" "LBM_TP_NS=6"; do
  env $cfg timeout 300 python bench.py --workload mrtcg_rt_weak $B 2>gpurun_out/c3_err.txt | tail -1 > gpurun_out/c3_mrtcg_${cfg}.json; show "mrtcg 8192x16384 $cfg" gpurun_out/c3_mrtcg_${cfg}.json
done
env LBM_TP_NS=5 timeout 300 python bench.py --workload mrtcg_rt $B 2>>gpurun_out/c3_err.txt | tail -1 > gpurun_out/c3_mrtcg16k_NS5.json; show "mrtcg 16384^2 NS=5" gpurun_out/c3_mrtcg16k_NS5.json
for cfg in "LBM_TP_STAGED=0" "LBM_TP_NS=3" "LBM_TP_NS=4" "LBM_TP_NS=5" "LBM_TP_NS=6"; do
  env $cfg timeout 300 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c3_err.txt | tail -1 > gpurun_out/c3_rk_${cfg}.json; show "rk 4096^2 $cfg" gpurun_out/c3_rk_${cfg}.json
done
for cfg in "LBM_CSF_FUSED=0" "LBM_CSF_FUSED=1"; do
  env $cfg timeout 300 python bench.py --workload csf_rt $B 2>>gpurun_out/c3_err.txt | tail -1 > gpurun_out/c3_csf_${cfg}.json; show "csf 8192^2 $cfg" gpurun_out/c3_csf_${cfg}.json
  env $cfg timeout 300 python bench.py --lib lattice-boltzmann-method_b200/liblbm_b200_hints.so --workload csf_rt $B 2>>gpurun_out/c3_err.txt | tail -1 > gpurun_out/c3_csf_hints_${cfg}.json; show "csf 8192^2 hints=2 $cfg" gpurun_out/c3_csf_hints_${cfg}.json
done
env LBM_TP_STAGED=0 timeout 300 python bench.py --lib lattice-boltzmann-method_b200/liblbm_b200_hints.so --workload mrtcg_rt_weak $B 2>>gpurun_out/c3_err.txt | tail -1 > gpurun_out/c3_mrtcg_hints.json; show "mrtcg k_tp_fused hints=2" gpurun_out/c3_mrtcg_hints.json
tail -5 gpurun_out/c3_err.txt
CMD="python bench.py --workload mrtcg_rt_weak --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/c3_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tp_staged -s 2 -c 1 -o gpurun_out/r02_ncu_mrtcg_staged -f $CMD > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log | cut -c1-300
