mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/c2_tests.txt
cat gpurun_out/c2_tests.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench_full.json 2> gpurun_out/c2_bench_full.err ) 2>&1 | grep real
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c2_bench_ref.json 2>gpurun_out/c2_bench_ref.err
for g in on off; do for w in rk_droplet poiseuille; do
  timeout 200 python bench.py --workload $w --graph $g --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/c2_${w}_graph_$g.json
done; done
python - <<'PY'
import json,glob
j=json.load(open('gpurun_out/c2_bench_full.json'))
print('headline %.2f GLUPS frac %.3f e2e %.2f blocks %d' % (j['value']/1e3, j['roofline']['frac'], j['e2e']['value']/1e3, j['timing']['blocks']), j['cpu_baseline']['value'], j['cpu_baseline']['cores'])
for k,v in (j['other_workloads'] or {}).items():
    print('%-18s %8.2f GLUPS  kernel %.3f  step %.3f  graph %s  %.1fs' % (k, v['value']/1e3, v['roofline']['frac'] or 0, v['roofline']['whole_step_frac_per_gpu'], v['cuda_graph'], v['setup_and_run_seconds']))
for f in sorted(glob.glob('gpurun_out/c2_*_graph_*.json')):
    try:
        v=json.load(open(f)); print(f, '%.2f GLUPS kernel %.3f step %.3f' % (v['value']/1e3, v['roofline']['frac'], v['roofline']['whole_step_frac_per_gpu']))
    except Exception as e: print(f, 'FAILED', e)
print(open('gpurun_out/c2_bench_ref.json').read()[:300])
PY
# ncu: single-pass CSF kernel and the ADE instantiation (4096 x 8192), only after the plain run of the same command exited 0
CMD="python bench.py --workload csf_rt --X 4096 --Y 8192 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
LBM_CSF_FUSED=1 $CMD > gpurun_out/c2_plain_csf.log 2>&1 && LBM_CSF_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_csf_fused -s 2 -c 1 -o gpurun_out/r02_ncu_csf_fused -f $CMD > gpurun_out/c2_ncu_csf.log 2>&1
CMD2="python bench.py --workload sedimentation --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/c2_plain_sed.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bgk_interior -s 7 -c 1 -o gpurun_out/r02_ncu_sedimentation -f $CMD2 > gpurun_out/c2_ncu_sed.log 2>&1
ls -la gpurun_out/*.ncu-rep
