set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline --no-check "$@" > gpurun_out/r2p_bench_$name.json 2> gpurun_out/r2p_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2p_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2p_bench_$name.err; }
use() { cp raglite_b200/lib/ab_$1.so raglite_b200/lib/libraglite_b200.so; touch raglite_b200/lib/libraglite_b200.so; }
for rep in 1 2 3; do
use old;  run old_c4_$rep --steps 30 --warmup 5
use sign; run sign_c4_$rep --steps 30 --warmup 5
done
for rep in 1 2; do
use old;  run old_fp16_$rep --steps 30 --warmup 5 --storage fp16
use sign; run sign_fp16_$rep --steps 30 --warmup 5 --storage fp16
done
use sign
date
