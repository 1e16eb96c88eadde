set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
date
# 1. the unverified fp16 tensor-map path first, on its own: a failure here flips the rest of the run to RL_TC_TMA=0
timeout 400 python -m pytest tests/test_gpu_search.py -m gpu -q -k "fp16 or storage or inserts" --timeout 300 -p no:cacheprovider > gpurun_out/r2s_pytest_fp16.log 2>&1
rc16=$?; echo "pytest fp16 rc=$rc16"; tail -5 gpurun_out/r2s_pytest_fp16.log
if [ $rc16 -ne 0 ]; then export RL_TC_TMA=0; echo "TMA path FAILED: continuing with RL_TC_TMA=0"; fi
# 2. the whole GPU suite
rm -f gpurun_out/scale_parity.jsonl
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2s_pytest_gpu.log
date
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2s_bench_$name.json 2> gpurun_out/r2s_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d.get("stage_ms",{}).get("main_scan",0),3), d.get("check",{}).get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2s_bench_$name.err; }
run c4 --steps 20 --warmup 3
run c4_fp16 --steps 20 --warmup 3 --storage fp16
RL_TC_TMA=0 run c4_fp16_reg --steps 20 --warmup 3 --storage fp16
run c3 --workload c3 --steps 10 --warmup 3
date
# 3. launch list of the default bench command + one full capture of the dominant kernel as it stands (cta_group::2 pairs)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-check > gpurun_out/r2s_launches_c4.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2s_ncu_pair python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/r2s_ncu_pair.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2s_ncu_pair.ncu-rep --page raw --csv > gpurun_out/r2s_ncu_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/r2s_ncu_pair.ncu-rep --page source --csv > gpurun_out/r2s_ncu_pair_source.csv 2>/dev/null
ls -la gpurun_out/ | grep r2s_
date
