set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
date
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2zz_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2zz_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2zz_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2zz_smoke.log
timeout 300 python bench.py > gpurun_out/r2zz_bench_default.json 2> gpurun_out/r2zz_bench_default.err; echo "default rc=$?"; cut -c1-200 gpurun_out/r2zz_bench_default.json
date
