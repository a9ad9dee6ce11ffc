mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== full bench"; ( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c10_bench_full.json 2> gpurun_out/c10_bench_full.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.load(open('gpurun_out/c10_bench_full.json'))
print('headline %.2f GLUPS frac %.3f e2e %.2f blocks %d' % (j['value']/1e3, j['roofline']['frac'], j['e2e']['value']/1e3, j['timing']['blocks']), j['cpu_baseline']['value'], j['cpu_baseline']['cores'], j['clocks'])
for k,v in (j['other_workloads'] or {}).items():
    print('%-18s %8.2f GLUPS  kernel %.3f  step %.3f  graph %s  %.1fs' % (k, v['value']/1e3, v['roofline']['frac'] or 0, v['roofline']['whole_step_frac_per_gpu'], v['cuda_graph'], v['setup_and_run_seconds']))
PY
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[2])); r=j['roofline']
    print('%-40s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f' % (sys.argv[1], j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
for rpb in 0 24 32 40 48 58 67 84 128; do
  LBM_TP_RPB=$rpb timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c10_err.txt | tail -1 > gpurun_out/c10_rk_$rpb.json; show "rk 4096^2 stash NS=4 rpb=$rpb" gpurun_out/c10_rk_$rpb.json
done
for rpb in 32 48 62 103; do
  LBM_TP_NS=3 LBM_TP_RPB=$rpb timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c10_err.txt | tail -1 > gpurun_out/c10_rk3_$rpb.json; show "rk 4096^2 stash NS=3 rpb=$rpb" gpurun_out/c10_rk3_$rpb.json
done
for rpb in 32 58 128; do
  LBM_TP_STASH=0 LBM_TP_RPB=$rpb timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/c10_err.txt | tail -1 > gpurun_out/c10_rk0_$rpb.json; show "rk 4096^2 STASH=0 rpb=$rpb" gpurun_out/c10_rk0_$rpb.json
done
tail -2 gpurun_out/c10_err.txt
