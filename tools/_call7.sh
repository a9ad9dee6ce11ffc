mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-44s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f  launches/step %.1f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu'], j['gpu_launches']/j['steps']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
i=0
for cfg in "A=1" "LBM_TP_STASH=0" "LBM_TP_NS=3" "LBM_TP_STASH=0 LBM_TP_NS=4" "LBM_TP_STASH=0 LBM_TP_NS=6"; do i=$((i+1))
  env $cfg timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c7_err.txt | tail -1 > gpurun_out/c7_rk_$i.json; show "rk 4096^2 $cfg" gpurun_out/c7_rk_$i.json
done
i=0
for cfg in "A=1" "LBM_TP_STASH=0"; do i=$((i+1))
  env $cfg timeout 200 python bench.py --workload mrtcg_rt_weak $B 2>gpurun_out/c7_err.txt | tail -1 > gpurun_out/c7_mrtcg_$i.json; show "mrtcg 8192x16384 $cfg" gpurun_out/c7_mrtcg_$i.json
done
timeout 300 python bench.py --workload mrtcg_rt $B 2>>gpurun_out/c7_err.txt | tail -1 > gpurun_out/c7_mrtcg16k.json; show "mrtcg 16384^2 default" gpurun_out/c7_mrtcg16k.json
tail -3 gpurun_out/c7_err.txt
CMD="python bench.py --workload kbc_shear --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/c7_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgk_interior -s 7 -c 1 -o gpurun_out/r02_ncu_kbc -f $CMD > gpurun_out/c7_ncu.log 2>&1
CMD2="python bench.py --workload poiseuille --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --graph off"
$CMD2 > gpurun_out/c7_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgk_interior -s 7 -c 1 -o gpurun_out/r02_ncu_poiseuille -f $CMD2 > gpurun_out/c7_ncu2.log 2>&1
ls -la gpurun_out/r02_ncu_kbc.ncu-rep gpurun_out/r02_ncu_poiseuille.ncu-rep
