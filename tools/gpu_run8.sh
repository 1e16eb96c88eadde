set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_search.py -m gpu -q --maxfail=10 --timeout 600 -p no:cacheprovider -k "rerank or xenc or linear or selection_larger or merge_of_more" > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2h_pytest_gpu.log
timeout 300 python tools/time_linear.py > gpurun_out/r2h_time_linear.json 2> gpurun_out/r2h_time_linear.err; echo "tl rc=$?"; cat gpurun_out/r2h_time_linear.json; tail -3 gpurun_out/r2h_time_linear.err
timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2h_bench_c5_mc.json 2> gpurun_out/r2h_bench_c5_mc.err; echo "c5 rc=$?"; head -c 300 gpurun_out/r2h_bench_c5_mc.json; tail -2 gpurun_out/r2h_bench_c5_mc.err
RL_XENC_MC=0 timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2h_bench_c5_nomc.json 2> gpurun_out/r2h_bench_c5_nomc.err; echo "c5 rc=$?"; head -c 300 gpurun_out/r2h_bench_c5_nomc.json
