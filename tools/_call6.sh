# usage: _call6.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | head -8
if [ "$N" = 2 ]; then
  echo "== pytest test_gpu_multi"; timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
  echo "== mp_nccl_check extra"; LBM_RING_EXTRA=1 timeout 300 $T --master-port 29514 tests/mp_nccl_check.py 2>&1 | grep -E "ring|rror|Traceback" | tee gpurun_out/c6_mp_nccl_check_n$N.log
fi
echo "== bench --gpus $N"; ( time timeout 600 $T --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c6_bench_n$N.json 2> gpurun_out/c6_bench_n$N.err ) 2>&1 | grep real
tail -3 gpurun_out/c6_bench_n$N.err
echo "== reference arm under torchrun"; timeout 300 $T --master-port 29513 bench.py --impl reference --gpus $N --steps 10 --warmup 2 2>/dev/null | grep "^{" > gpurun_out/c6_ref_n$N.json
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.loads([l for l in open(f'gpurun_out/c6_bench_n{n}.json') if l.startswith('{')][-1])
    print('headline %.2f GLUPS (%.3f ms/step) frac %.3f e2e %.2f' % (j['value']/1e3, j['ms_per_step'], j['roofline']['frac'], j['e2e']['value']/1e3))
    print('ring_parity', j['ring_parity'])
    for k,v in (j['other_workloads'] or {}).items(): print(k, '%.2f GLUPS  kernel %.3f step %.3f' % (v['value']/1e3, v['roofline']['frac'], v['roofline']['whole_step_frac_per_gpu']))
except Exception as e: print('bench FAILED', e)
try:
    r=json.load(open(f'gpurun_out/c6_ref_n{n}.json')); print('reference arm', r['value'], r['cpu_baseline'])
except Exception as e: print('ref FAILED', e)
PY
