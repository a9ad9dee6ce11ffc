mkdir -p gpurun_out
echo "== csf tests"; timeout 900 python -m pytest tests/test_gpu_csf.py tests/test_gpu_bench_scale.py::test_csf_at_8192 -m gpu -q 2>&1 | tail -2
tools/gpu_checks.sh ab 2>&1 | grep csf
tools/gpu_checks.sh ncu csf_rt k_csf_staged 2>&1 | tail -1
