set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
date
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2t_bench_$name.json 2> gpurun_out/r2t_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2t_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d.get("stage_ms",{}).get("main_scan",0),3), d.get("check",{}).get("identical_topk_sets"), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2t_bench_$name.err; }
# configs[2] (B = 1024): pairs were measured with the slow generic loader; again with the fast one, both storages
run c3 --workload c3 --steps 10 --warmup 3
RL_TC_PAIR=1 run c3_pair --workload c3 --steps 10 --warmup 3
run c3_b --workload c3 --steps 10 --warmup 3
RL_TC_PAIR=1 run c3_pair_b --workload c3 --steps 10 --warmup 3
run c3_fp16 --workload c3 --steps 10 --warmup 3 --storage fp16
RL_TC_PAIR=1 run c3_fp16_pair --workload c3 --steps 10 --warmup 3 --storage fp16
date
timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2t_ncu_c3 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/r2t_ncu_c3.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2t_ncu_c3.ncu-rep --page raw --csv > gpurun_out/r2t_ncu_c3_raw.csv 2>/dev/null
ncu -i gpurun_out/r2t_ncu_c3.ncu-rep --page source --csv > gpurun_out/r2t_ncu_c3_source.csv 2>/dev/null
RL_TC_PAIR=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2t_ncu_c3_pair python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/r2t_ncu_c3_pair.log 2>&1; echo "ncu pair rc=$?"
ncu -i gpurun_out/r2t_ncu_c3_pair.ncu-rep --page raw --csv > gpurun_out/r2t_ncu_c3_pair_raw.csv 2>/dev/null
rm -f gpurun_out/r2t_ncu_c3_pair.ncu-rep
date
