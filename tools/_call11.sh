mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-40s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
echo "== two-phase tests"; timeout 900 python -m pytest tests/test_gpu_two_phase.py tests/test_gpu_csf.py tests/test_gpu_slabs.py tests/test_gpu_bench_scale.py -m gpu -q 2>&1 | tail -2
for w in rk_droplet mrtcg_rt_weak csf_rt; do
  s=10; [ $w = rk_droplet ] && s=50
  timeout 200 python bench.py --workload $w --steps $s --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c11_err.txt | tail -1 > gpurun_out/c11_$w.json; show "$w default" gpurun_out/c11_$w.json
done
LBM_TP_STASH=1 timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c11_err.txt | tail -1 > gpurun_out/c11_rk_stash.json; show "rk stash" gpurun_out/c11_rk_stash.json
CMD="python bench.py --workload poiseuille --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --graph off"
$CMD > gpurun_out/c11_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/r02_ncu_launches_poiseuille.csv $CMD > gpurun_out/c11_ncu.log 2>&1
CMD2="python bench.py --workload rk_droplet --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --graph off"
$CMD2 > gpurun_out/c11_plain2.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 64 --csv --log-file gpurun_out/r02_ncu_launches_rk.csv $CMD2 > gpurun_out/c11_ncu2.log 2>&1
python tools/ncu_launch_list.py gpurun_out/r02_ncu_launches_poiseuille.csv "poiseuille 2700x2100" 2>&1 | head -30
python tools/ncu_launch_list.py gpurun_out/r02_ncu_launches_rk.csv "rk 4096^2" 2>&1 | head -30
