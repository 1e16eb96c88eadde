set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline --no-check "$@" > gpurun_out/r2o_bench_$name.json 2> gpurun_out/r2o_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2o_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2o_bench_$name.err; }
nvidia-smi --query-gpu=power.limit,power.draw,clocks.sm,clocks.mem,temperature.gpu --format=csv
run sign_c4 --steps 10 --warmup 3
run sign_c4_fp16 --steps 10 --warmup 3 --storage fp16
RL_NVCC_EXTRA="-DRL_EPI_SIGN=0" python -c "from raglite_b200 import _build; _build.build(force=True); print('rebuilt old epilogue')"
run old_c4 --steps 10 --warmup 3
run old_c4_fp16 --steps 10 --warmup 3 --storage fp16
python -c "from raglite_b200 import _build; _build.build(force=True); print('rebuilt sign epilogue')"
run sign2_c4 --steps 10 --warmup 3
run sign2_c4_fp16 --steps 10 --warmup 3 --storage fp16
nvidia-smi --query-gpu=power.limit,power.draw,clocks.sm,clocks.mem,temperature.gpu --format=csv
date
