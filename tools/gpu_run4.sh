set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_search.py -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider -k "rerank or xenc or rank_then_filter or metadata" > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2d_pytest_gpu.log
for b in 128 192; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/r2d_bench_c4_b$b.json 2> gpurun_out/r2d_bench_c4_b$b.err; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --filtered > gpurun_out/r2d_bench_c4_filtered.json 2> gpurun_out/r2d_bench_c4_filtered.err; echo "c4f rc=$?"
for f in c4_b128 c4_b192 c4_filtered; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2d_bench_$f.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), d["stage_ms"]["main_scan"], d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3), d.get("filtered"))
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/r2d_bench_$f.err; done
timeout 300 python tools/time_linear.py > gpurun_out/r2d_time_linear.json 2> gpurun_out/r2d_time_linear.err; echo "tl rc=$?"; cat gpurun_out/r2d_time_linear.json; tail -3 gpurun_out/r2d_time_linear.err
timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2d_bench_c5.json 2> gpurun_out/r2d_bench_c5.err; echo "c5 rc=$?"; head -c 300 gpurun_out/r2d_bench_c5.json; tail -2 gpurun_out/r2d_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear|attention|add_ln|cls_head|embed_ln" -c 120 --csv --log-file gpurun_out/r2d_launches_xenc.csv python tools/bench_rerank.py --pairs 700 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r2d_launches_xenc.log 2>&1; echo "ncu list rc=$?"
