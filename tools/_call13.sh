mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== full bench"; ( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c13_bench_full.json 2> gpurun_out/c13_bench_full.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.load(open('gpurun_out/c13_bench_full.json'))
e=j['e2e']
print('headline %.2f GLUPS frac %.3f step %.3f e2e %.2f (blocking %.2f) blocks %d' % (j['value']/1e3, j['roofline']['frac'], j['roofline']['whole_step_frac_per_gpu'], e['value']/1e3, e['blocking']['value']/1e3, j['timing']['blocks']), j['cpu_baseline']['value'], j['cpu_baseline']['cores'], j['clocks'])
print(j['roofline']['other_spans_ms_per_step'], j['roofline']['kernel_ms_per_step'], j['ms_per_step'])
for k,v in (j['other_workloads'] or {}).items():
    print('%-18s %8.2f GLUPS  kernel %.3f  step %.3f  graph %s  %.1fs' % (k, v['value']/1e3, v['roofline']['frac'] or 0, v['roofline']['whole_step_frac_per_gpu'], v['cuda_graph'], v['setup_and_run_seconds']))
PY
for g in on off; do timeout 200 python bench.py --workload poiseuille --graph $g --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; print('poiseuille graph $g %.2f GLUPS kernel %.3f (%.4f ms) step %.3f' % (j['value']/1e3, r['frac'], r['kernel_ms_per_step'], r['whole_step_frac_per_gpu']), r['other_spans_ms_per_step'])"; done
