mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-46s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
echo "== two-phase tests"; timeout 900 python -m pytest tests/test_gpu_two_phase.py tests/test_gpu_csf.py tests/test_gpu_slabs.py tests/test_gpu_bench_scale.py -m gpu -q 2>&1 | tail -3
echo "== csf staged tests"; LBM_CSF_FUSED=1 LBM_CSF_STAGED=2 timeout 600 python -m pytest tests/test_gpu_csf.py tests/test_gpu_bench_scale.py::test_csf_at_8192 -m gpu -q 2>&1 | tail -2
i=0
for cfg in "A=1" "LBM_TP_RPB=128" "LBM_TP_STASH=0" "LBM_TP_STASH=0 LBM_TP_RPB=128" "LBM_TP_NS=2"; do i=$((i+1))
  env $cfg timeout 200 python bench.py --workload mrtcg_rt_weak $B 2>gpurun_out/c8_err.txt | tail -1 > gpurun_out/c8_mrtcg_$i.json; show "mrtcg 8192x16384 $cfg" gpurun_out/c8_mrtcg_$i.json
done
timeout 300 python bench.py --workload mrtcg_rt $B 2>>gpurun_out/c8_err.txt | tail -1 > gpurun_out/c8_mrtcg16k.json; show "mrtcg 16384^2 default" gpurun_out/c8_mrtcg16k.json
i=0
for cfg in "A=1" "LBM_TP_STASH=0" "LBM_TP_NS=3" "LBM_TP_NS=2"; do i=$((i+1))
  env $cfg timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c8_err.txt | tail -1 > gpurun_out/c8_rk_$i.json; show "rk 4096^2 $cfg" gpurun_out/c8_rk_$i.json
done
i=0
for cfg in "LBM_CSF_FUSED=1" "LBM_CSF_FUSED=1 LBM_CSF_STAGED=2"; do i=$((i+1))
  env $cfg timeout 300 python bench.py --workload csf_rt $B 2>>gpurun_out/c8_err.txt | tail -1 > gpurun_out/c8_csf_$i.json; show "csf 8192^2 $cfg" gpurun_out/c8_csf_$i.json
done
tail -3 gpurun_out/c8_err.txt
