set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_scale.py -m gpu -q --maxfail=10 --timeout 600 -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2m_pytest.log
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2m_bench_$name.json 2> gpurun_out/r2m_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2m_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2m_bench_$name.err; }
run c4 --steps 10 --warmup 3
run c4_b --steps 10 --warmup 3
RL_TC_PAIR=0 run c4_single --steps 10 --warmup 3
run c3 --workload c3 --steps 5 --warmup 3
run c2 --workload c2 --steps 20 --warmup 3
./tools/membw_probe > gpurun_out/r2m_membw.txt 2>&1; cat gpurun_out/r2m_membw.txt
date
