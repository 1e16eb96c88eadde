cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_rerank.py -m gpu -q --timeout 50 -p no:cacheprovider > gpurun_out/r3c_pytest_rerank.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3c_pytest_rerank.log
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention" -c 24 --csv --log-file gpurun_out/r3c_launches_att_async.csv python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r3c_bench_rerank.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r3c_bench_rerank.log | cut -c1-400
RL_XENC_ATT_CPASYNC=0 timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention" -c 24 --csv --log-file gpurun_out/r3c_launches_att_sync.csv python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > /dev/null 2>&1; echo "rc=$?"
