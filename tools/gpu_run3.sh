set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider -k "several_query_groups or c2_full" > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2c_pytest_gpu.log
timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c3_par.json 2> gpurun_out/r2c_bench_c3_par.err; echo "c3 rc=$?"
RL_TC_PAIR=1 timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c3_par_pair.json 2> gpurun_out/r2c_bench_c3_par_pair.err; echo "c3pp rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --filtered > gpurun_out/r2c_bench_c4_filtered.json 2> gpurun_out/r2c_bench_c4_filtered.err; echo "c4f rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --storage fp16 > gpurun_out/r2c_bench_c4_fp16.json 2> gpurun_out/r2c_bench_c4_fp16.err; echo "c4h rc=$?"
for f in c3_par c3_par_pair c4_filtered c4_fp16; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_bench_$f.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), d["stage_ms"], d["check"].get("identical_topk_sets"), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3), d.get("filtered"))
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/r2c_bench_$f.err; done
timeout 300 python tools/time_linear.py > gpurun_out/r2c_time_linear.json 2> gpurun_out/r2c_time_linear.err; echo "tl rc=$?"; cat gpurun_out/r2c_time_linear.json; tail -3 gpurun_out/r2c_time_linear.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear|attention|add_ln|cls_head|embed_ln" -c 200 --csv --log-file gpurun_out/r2c_launches_xenc.csv python tools/bench_rerank.py --pairs 700 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r2c_launches_xenc.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"linear_wres|attention" -s 10 -c 4 -f -o gpurun_out/r2_ncu_xenc python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r2_ncu_xenc.log 2>&1; echo "ncu xenc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"linear_tcgen05" -s 3 -c 1 -f -o gpurun_out/r2_ncu_xenc_down python tools/bench_rerank.py --pairs 300 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r2_ncu_xenc_down.log 2>&1; echo "ncu xenc2 rc=$?"
ls -la gpurun_out/*.ncu-rep
