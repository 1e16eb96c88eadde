set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/scale_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2e_pytest_gpu.log
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2e_bench_$name.json 2> gpurun_out/r2e_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2e_bench_$name.err; }
run c4_dual --steps 10 --warmup 3
RL_TC_DUAL=0 run c4_single --steps 10 --warmup 3
run c4_dual_b --steps 10 --warmup 3
RL_TC_DUAL=0 run c4_single_b --steps 10 --warmup 3
run c4_fp16_dual --steps 10 --warmup 3 --storage fp16
RL_TC_DUAL=0 run c4_fp16_single --steps 10 --warmup 3 --storage fp16
run c3_dual --workload c3 --steps 5 --warmup 3
RL_TC_DUAL=0 run c3_single --workload c3 --steps 5 --warmup 3
run c2_dual --workload c2 --steps 20 --warmup 3
RL_TC_DUAL=0 run c2_single --workload c2 --steps 20 --warmup 3
run c4_clustered_dual --steps 10 --warmup 3 --data clustered
timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err; echo "c5 rc=$?"; head -c 300 gpurun_out/r2e_bench_c5.json; tail -2 gpurun_out/r2e_bench_c5.err
