mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== full bench"; ( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c14_bench_full.json 2> gpurun_out/c14_bench_full.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.load(open('gpurun_out/c14_bench_full.json'))
e=j['e2e']
print('headline %.2f GLUPS frac %.3f step %.3f e2e %.2f (blocking %.2f) blocks %d' % (j['value']/1e3, j['roofline']['frac'], j['roofline']['whole_step_frac_per_gpu'], e['value']/1e3, e['blocking']['value']/1e3, j['timing']['blocks']), j['cpu_baseline']['value'], j['cpu_baseline']['cores'], j['clocks'])
for k,v in (j['other_workloads'] or {}).items():
    print('%-18s %8.2f GLUPS  kernel %.3f  step %.3f  graph %s  %.1fs' % (k, v['value']/1e3, v['roofline']['frac'] or 0, v['roofline']['whole_step_frac_per_gpu'], v['cuda_graph'], v['setup_and_run_seconds']))
PY
timeout 300 python bench.py --steps 100 --warmup 5 --no-others --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); e=j['e2e']; print('K=100: value %.2f e2e %.2f blocking %.2f' % (j['value']/1e3, e['value']/1e3, e['blocking']['value']/1e3))"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c14_bench_ref.json 2>/dev/null; head -c 300 gpurun_out/c14_bench_ref.json
