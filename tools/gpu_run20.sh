set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_scale.py -m gpu -q --maxfail=10 --timeout 600 -p no:cacheprovider > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2r_pytest.log
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2r_bench_$name.json 2> gpurun_out/r2r_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2r_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "serial", round(d["e2e"]["serial_ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -3 gpurun_out/r2r_bench_$name.err; }
run fp16_tma_1 --steps 20 --warmup 3 --storage fp16
RL_TC_TMA=0 run fp16_reg_1 --steps 20 --warmup 3 --storage fp16
run fp16_tma_2 --steps 20 --warmup 3 --storage fp16
RL_TC_TMA=0 run fp16_reg_2 --steps 20 --warmup 3 --storage fp16
run c4 --steps 20 --warmup 3
date
