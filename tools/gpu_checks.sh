#!/usr/bin/env bash
# One-call GPU check list (run under gpurun from the repo root).  Every command has its own short timeout: a hung
# multi-rank program is charged N x wall time, and one such hang cost this project the rest of a round's GPU budget.
#   tools/gpu_checks.sh tests          full -m gpu suite + smoke
#   tools/gpu_checks.sh bench          the default bench line (headline + other_workloads) and the reference arm
#   tools/gpu_checks.sh ab             same-box A/B lines of the two-phase kernels' switches
#   tools/gpu_checks.sh ring N         tests/mp_nccl_check.py (N = 2) and bench.py on N ranks (N GPUs requested)
#   tools/gpu_checks.sh ncu W KERNEL   ncu --set full of KERNEL (regex) in workload W -> gpurun_out/ncu_W.ncu-rep
#                                      (read it here with tools/ncu_summary.py)
# Boxes of the pool differ by up to 4 %: compare variants within one call only.
set -u
mkdir -p gpurun_out
what=${1:-tests}
B="--no-cpu-baseline --no-e2e"
line() {  # line LABEL FILE: one summary line of a bench JSON
  python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[2])); r = j["roofline"]
    print("%-46s %7.2f GLUPS  kernel %.3f (%.3f ms)  step %.3f" % (sys.argv[1], j["value"] / 1e3, r["frac"], r["kernel_ms_per_step"], r["whole_step_frac_per_gpu"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
case "$what" in
  tests)
    timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
    ;;
  bench)
    timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
    python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_full.json")); e = j["e2e"]
print("headline %.2f GLUPS  kernel %.3f  step %.3f  e2e %.2f (blocking %.2f)  cpu %.2f MLUPS on %d threads" % (
    j["value"] / 1e3, j["roofline"]["frac"], j["roofline"]["whole_step_frac_per_gpu"], e["value"] / 1e3, e["blocking"]["value"] / 1e3,
    j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"]))
for k, v in (j["other_workloads"] or {}).items():
    print("%-18s %8.2f GLUPS  kernel %.3f  step %.3f  graph %s" % (k, v["value"] / 1e3, v["roofline"]["frac"] or 0, v["roofline"]["whole_step_frac_per_gpu"], v["cuda_graph"]))
PY
    timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
    head -c 200 gpurun_out/bench_reference.json; echo
    ;;
  ab)
    i=0
    for cfg in "A=1" "LBM_TP_STASH=0" "LBM_TP_STAGED=0" "LBM_TP_NS=2" "LBM_TP_NS=4" "LBM_TP_RPB=128" "LBM_TP_RPB=32"; do i=$((i+1))
      env $cfg timeout 200 python bench.py --workload mrtcg_rt_weak --steps 10 --warmup 3 $B 2>/dev/null | tail -1 > gpurun_out/ab_mrtcg_$i.json; line "mrtcg 8192x16384 $cfg" gpurun_out/ab_mrtcg_$i.json
      env $cfg timeout 200 python bench.py --workload rk_droplet --steps 50 --warmup 5 $B 2>/dev/null | tail -1 > gpurun_out/ab_rk_$i.json; line "rk 4096^2 $cfg" gpurun_out/ab_rk_$i.json
    done
    i=0
    for cfg in "A=1" "LBM_CSF_STAGED=1" "LBM_CSF_STAGED=0" "LBM_CSF_STAGED=0 LBM_CSF_PIPE=1" "LBM_CSF_FUSED=0"; do i=$((i+1))
      env $cfg timeout 300 python bench.py --workload csf_rt --steps 10 --warmup 3 $B 2>/dev/null | tail -1 > gpurun_out/ab_csf_$i.json; line "csf 8192^2 $cfg" gpurun_out/ab_csf_$i.json
    done
    ;;
  ring)
    n=${2:-2}
    T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
    if [ "$n" = 2 ]; then
      LBM_RING_EXTRA=1 timeout 300 $T --master-port 29514 tests/mp_nccl_check.py 2>&1 | grep -E "ring|rror|Traceback" | tee gpurun_out/mp_nccl_check_n$n.log
      echo "mp_nccl_check rc=${PIPESTATUS[0]}"
    fi
    timeout 600 $T --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/bench_n$n.err | grep "^{" > gpurun_out/bench_n$n.json
    timeout 300 $T --master-port 29513 bench.py --impl reference --gpus $n --steps 10 --warmup 2 2>/dev/null | grep "^{" > gpurun_out/bench_reference_n$n.json
    python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.load(open(f"gpurun_out/bench_n{n}.json"))
    print("n=%s headline %.2f GLUPS (%.3f ms/step)  kernel %.3f  e2e %.2f" % (n, j["value"] / 1e3, j["ms_per_step"], j["roofline"]["frac"], j["e2e"]["value"] / 1e3))
    print("ring_parity", j["ring_parity"])
    for k, v in (j["other_workloads"] or {}).items():
        print(k, "%.2f GLUPS  kernel %.3f  step %.3f" % (v["value"] / 1e3, v["roofline"]["frac"], v["roofline"]["whole_step_frac_per_gpu"]))
    r = json.load(open(f"gpurun_out/bench_reference_n{n}.json"))
    print("reference arm under torchrun: %.2f MLUPS on %d threads" % (r["value"], r["cpu_baseline"]["cores"]))
except Exception as e:
    print("FAILED", e)
PY
    ;;
  ncu)
    w=${2:?workload}; k=${3:?kernel regex}
    CMD="python bench.py --workload $w --steps 3 --warmup 3 $B --graph off"
    $CMD > gpurun_out/ncu_plain_$w.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 \
        -o gpurun_out/ncu_$w -f $CMD > gpurun_out/ncu_$w.log 2>&1
    ls -la gpurun_out/ncu_$w.ncu-rep
    ;;
  *) echo "usage: tools/gpu_checks.sh tests|bench|ab|ring N|ncu W KERNEL"; exit 2 ;;
esac
