set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/scale_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2j_pytest_gpu.log
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2j_bench_$name.json 2> gpurun_out/r2j_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2j_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2j_bench_$name.err; }
run c4 --steps 10 --warmup 3
run c4_b --steps 10 --warmup 3
run c4_fp16 --steps 10 --warmup 3 --storage fp16
run c3 --workload c3 --steps 5 --warmup 3
run c2 --workload c2 --steps 20 --warmup 3
run c4_clustered --steps 10 --warmup 3 --data clustered
date
