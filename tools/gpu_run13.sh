set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/scale_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2l_pytest_gpu.log
run() { name=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2l_bench_$name.json 2> gpurun_out/r2l_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "scan", round(d["stage_ms"]["main_scan"],3), d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("$name ERR", e)
PY
tail -2 gpurun_out/r2l_bench_$name.err; }
run c4 --steps 10 --warmup 3
RL_TC_PAIR=0 run c4_single --steps 10 --warmup 3
run c4_fp16 --steps 10 --warmup 3 --storage fp16
run c2 --workload c2 --steps 20 --warmup 3
RL_TC_PAIR=0 run c2_single --workload c2 --steps 20 --warmup 3
run c4_clustered --steps 10 --warmup 3 --data clustered
run c4_b128 --steps 10 --warmup 3 --batch 128
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2l_ncu_pair python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/r2l_ncu_pair.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2l_ncu_pair.ncu-rep --page raw --csv > gpurun_out/r2l_ncu_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/r2l_ncu_pair.ncu-rep --page source --csv > gpurun_out/r2l_ncu_pair_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2l_ncu_pair16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check --storage fp16 > gpurun_out/r2l_ncu_pair16.log 2>&1; echo "ncu16 rc=$?"
ncu -i gpurun_out/r2l_ncu_pair16.ncu-rep --page raw --csv > gpurun_out/r2l_ncu_pair16_raw.csv 2>/dev/null
ncu -i gpurun_out/r2l_ncu_pair16.ncu-rep --page source --csv > gpurun_out/r2l_ncu_pair16_source.csv 2>/dev/null
ls -la gpurun_out/ | grep r2l_ncu
date
