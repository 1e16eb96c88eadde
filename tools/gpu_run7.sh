set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adapter_fit.py tests/test_gpu_search.py tests/test_fusion.py -m gpu -q --maxfail=10 --timeout 600 -p no:cacheprovider > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2g_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
