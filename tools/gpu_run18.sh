set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp raglite_b200/lib/ab_sign.so raglite_b200/lib/libraglite_b200.so; touch raglite_b200/lib/libraglite_b200.so
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -q --maxfail=10 --timeout 600 -p no:cacheprovider -k "async or dropin or overflow" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2q_pytest.log
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2q_bench_$name.json 2> gpurun_out/r2q_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2q_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "serial", round(d["e2e"]["serial_ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -3 gpurun_out/r2q_bench_$name.err; }
run c4 --steps 20 --warmup 3
run c4_if3 --steps 20 --warmup 3 --inflight 3
run c2 --workload c2 --steps 20 --warmup 3
run c3 --workload c3 --steps 5 --warmup 3
date
