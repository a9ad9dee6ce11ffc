"""CPU: world_size-2 runs over gloo of the host-side multi-GPU logic, and of bench.py's reference arm
under torchrun (rank 0 alone works and prints; the other rank exits 0 without output)."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun(script_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port())] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


def test_host_side_slab_logic_world_size_2():
    r = torchrun([os.path.join("tests", "mp_gloo_check.py")])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_reference_arm_under_torchrun_prints_once():
    r = torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--cpu-sample", "96"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["metric"] == "MLUPS" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
