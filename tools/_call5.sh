mkdir -p gpurun_out
echo "== tests (defaults: staged + stash)"; timeout 1200 python -m pytest tests/test_gpu_two_phase.py tests/test_gpu_slabs.py tests/test_gpu_csf.py tests/test_gpu_long_horizon.py tests/test_gpu_graph.py tests/test_gpu_bench_scale.py -m gpu -q 2>&1 | tail -4
echo "== csf tests, staged + stash"; LBM_CSF_FUSED=1 LBM_CSF_STAGED=2 timeout 600 python -m pytest tests/test_gpu_csf.py tests/test_gpu_bench_scale.py::test_csf_at_8192 -m gpu -q 2>&1 | tail -3
echo "== csf tests, staged"; LBM_CSF_FUSED=1 LBM_CSF_STAGED=1 timeout 600 python -m pytest tests/test_gpu_csf.py -m gpu -q 2>&1 | tail -3
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-44s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
i=0
for cfg in "A=1" "LBM_TP_NS=2" "LBM_TP_NS=4" "LBM_TP_STASH=0"; do i=$((i+1))
  env $cfg timeout 200 python bench.py --workload mrtcg_rt_weak $B 2>gpurun_out/c5_err.txt | tail -1 > gpurun_out/c5_mrtcg_$i.json; show "mrtcg 8192x16384 $cfg" gpurun_out/c5_mrtcg_$i.json
  env $cfg timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c5_err.txt | tail -1 > gpurun_out/c5_rk_$i.json; show "rk 4096^2 $cfg" gpurun_out/c5_rk_$i.json
done
env LBM_TP_NS=3 timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c5_err.txt | tail -1 > gpurun_out/c5_rk_ns3.json; show "rk 4096^2 NS=3" gpurun_out/c5_rk_ns3.json
timeout 300 python bench.py --workload mrtcg_rt $B 2>>gpurun_out/c5_err.txt | tail -1 > gpurun_out/c5_mrtcg16k.json; show "mrtcg 16384^2 default" gpurun_out/c5_mrtcg16k.json
i=0
for cfg in "LBM_CSF_FUSED=0" "LBM_CSF_FUSED=1" "LBM_CSF_FUSED=1 LBM_CSF_STAGED=1" "LBM_CSF_FUSED=1 LBM_CSF_STAGED=2"; do i=$((i+1))
  env $cfg timeout 300 python bench.py --workload csf_rt $B 2>>gpurun_out/c5_err.txt | tail -1 > gpurun_out/c5_csf_$i.json; show "csf 8192^2 $cfg" gpurun_out/c5_csf_$i.json
done
tail -3 gpurun_out/c5_err.txt
CMD="python bench.py --workload csf_rt --X 4096 --Y 8192 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
LBM_CSF_FUSED=1 LBM_CSF_STAGED=2 $CMD > gpurun_out/c5_plain.log 2>&1 && LBM_CSF_FUSED=1 LBM_CSF_STAGED=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_csf_staged -s 2 -c 1 -o gpurun_out/r02_ncu_csf_staged -f $CMD > gpurun_out/c5_ncu.log 2>&1
CMD2="python bench.py --workload rk_droplet --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --graph off"
$CMD2 > gpurun_out/c5_plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tp_staged -s 2 -c 1 -o gpurun_out/r02_ncu_rk_stash -f $CMD2 > gpurun_out/c5_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
