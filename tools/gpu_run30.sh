set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
date
timeout 300 python -m pytest tests/test_gpu_rerank.py -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/r3a_pytest_rerank.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3a_pytest_rerank.log
RL_XENC_ATT_ORDER=1 timeout 300 python -m pytest tests/test_gpu_rerank.py -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/r3a_pytest_rerank_seqfast.log 2>&1; echo "pytest seqfast rc=$?"; tail -2 gpurun_out/r3a_pytest_rerank_seqfast.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention" -c 24 --csv --log-file gpurun_out/r3a_launches_att_headfast.csv python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > /dev/null 2>&1; echo "rc=$?"
RL_XENC_ATT_ORDER=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention" -c 24 --csv --log-file gpurun_out/r3a_launches_att_seqfast.csv python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > /dev/null 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention2|cls_head" -s 3 -c 2 -f -o gpurun_out/r3a_ncu_att python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r3a_ncu_att.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r3a_ncu_att.ncu-rep --page raw --csv > gpurun_out/r3a_ncu_att_raw.csv 2>/dev/null
rm -f gpurun_out/r3a_ncu_att.ncu-rep
date
