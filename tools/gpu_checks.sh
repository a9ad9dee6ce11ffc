#!/usr/bin/env bash
# One-call GPU check list (run under gpurun from the repo root).  Every command has its own short timeout: a hung
# multi-rank program is charged N x wall time, and one such hang cost this project the rest of a round's GPU budget.
#   tools/gpu_checks.sh tests          full -m gpu suite + smoke
#   tools/gpu_checks.sh bench          one bench line per workload into gpurun_out/bench_<workload>.json
#   tools/gpu_checks.sh ring N         tests/mp_nccl_check.py and the cylinder / mrtcg bench on N ranks (N GPUs requested)
set -u
mkdir -p gpurun_out
what=${1:-tests}
case "$what" in
  tests)
    timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
    ;;
  bench)
    for w in cylinder cylinder_bb kbc_shear poiseuille sedimentation sedimentation_ibm rk_droplet mrtcg_rt csf_rt; do
      s=100; [ "$w" = mrtcg_rt ] && s=20; [ "$w" = csf_rt ] && s=30
      timeout 300 python bench.py --workload $w --steps $s --warmup 5 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.json
      python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    j = json.load(open(f"gpurun_out/bench_{w}.json"))
    r = j["roofline"]
    print(f"{w:14s} {j['value'] / 1e3:7.2f} GLUPS  kernel frac {r['frac']:.3f}  whole step {r['whole_step_frac_per_gpu']:.3f}  e2e {j['e2e']['value'] / 1e3:.2f}")
except Exception as e:
    print(w, "FAILED", e)
PY
    done
    # A/B of the single-pass CSF step (off by default until this line says it is faster)
    LBM_CSF_FUSED=1 timeout 300 python bench.py --workload csf_rt --steps 30 --warmup 5 2> gpurun_out/bench_csf_rt_fused.err | tail -1 > gpurun_out/bench_csf_rt_fused.json
    python -c "import json; j=json.load(open('gpurun_out/bench_csf_rt_fused.json')); print('csf_rt fused   %7.2f GLUPS  kernel frac %.3f  whole step %.3f' % (j['value']/1e3, j['roofline']['frac'], j['roofline']['whole_step_frac_per_gpu']))" || echo "csf_rt fused FAILED"
    LBM_CSF_FUSED=1 LBM_CSF_PIPE=1 timeout 300 python bench.py --workload csf_rt --steps 30 --warmup 5 2> gpurun_out/bench_csf_rt_fused_pipe.err | tail -1 > gpurun_out/bench_csf_rt_fused_pipe.json
    python -c "import json; j=json.load(open('gpurun_out/bench_csf_rt_fused_pipe.json')); print('csf_rt fused+pipe %5.2f GLUPS  kernel frac %.3f  whole step %.3f' % (j['value']/1e3, j['roofline']['frac'], j['roofline']['whole_step_frac_per_gpu']))" || echo "csf_rt fused+pipe FAILED"
    ;;
  ring)
    n=${2:-2}
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
    timeout 240 $T --master-port 29511 tests/mp_nccl_check.py 2>&1 | grep -E "ring|rror|Traceback" | tee gpurun_out/mp_nccl_check_n$n.log
    echo "mp_nccl_check rc=${PIPESTATUS[0]}"
    timeout 180 $T --master-port 29512 bench.py --gpus $n --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_n${n}_cylinder.json
    timeout 240 $T --master-port 29513 bench.py --gpus $n --workload mrtcg_rt --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n${n}_mrtcg_rt.json
    python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
for w in ("cylinder", "mrtcg_rt"):
    try:
        j = json.load(open(f"gpurun_out/bench_n{n}_{w}.json"))
        print(f"n={n} {w:10s} {j['value'] / 1e3:8.2f} GLUPS  {j['ms_per_step']:.3f} ms/step")
    except Exception as e:
        print(w, "FAILED", e)
PY
    # last, under its own timeout: the checks whose all-to-all groups have only run over the NCCL stand-in so far
    # (lbm_comm_check, the RK diagnostics' ring-wide max) and the single-pass CSF step on the ring
    LBM_RING_EXTRA=1 LBM_CSF_FUSED=1 timeout 240 $T --master-port 29514 tests/mp_nccl_check.py 2>&1 | grep -E "ring|rror|Traceback" | tee gpurun_out/mp_nccl_check_extra_n$n.log
    echo "mp_nccl_check (extra) rc=${PIPESTATUS[0]}"
    ;;
  *) echo "usage: tools/gpu_checks.sh tests|bench|ring N"; exit 2 ;;
esac
