set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
date
timeout 400 python -m pytest tests/test_gpu_rerank.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2w_pytest_rerank.log 2>&1; echo "pytest rerank rc=$?"; tail -5 gpurun_out/r2w_pytest_rerank.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear|attention|add_ln|cls_head|embed_ln|seq_order" -c 200 --csv --log-file gpurun_out/r2w_launches_xenc.csv python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r2w_launches_xenc.log 2>&1; echo "launch list rc=$?"
RL_XENC_ATT_LPT=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention" -c 24 --csv --log-file gpurun_out/r2w_launches_xenc_nolpt.csv python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r2w_launches_xenc_nolpt.log 2>&1; echo "launch list nolpt rc=$?"
c5() { name=$1; timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2w_bench_$name.json 2> gpurun_out/r2w_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2w_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],1), "pairs/s", round(d["ms_per_step"],1), "ms", round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2w_bench_$name.err; }
c5 c5_a
c5 c5_b
date
