set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scale.py tests/test_gpu_search.py tests/test_gpu_rerank.py -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2b_pytest_gpu.log
timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c3_multi.json 2> gpurun_out/r2b_bench_c3_multi.err; echo "c3 rc=$?"
RL_TC_GROUPS=1 timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c3_g1.json 2> gpurun_out/r2b_bench_c3_g1.err; echo "c3g1 rc=$?"
RL_TC_PAIR=1 timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c3_multi_pair.json 2> gpurun_out/r2b_bench_c3_multi_pair.err; echo "c3pair rc=$?"
RL_TC_PAIR=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c4_pair.json 2> gpurun_out/r2b_bench_c4_pair.err; echo "c4pair rc=$?"
timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2b_bench_c5_res.json 2> gpurun_out/r2b_bench_c5_res.err; echo "c5 rc=$?"
RL_XENC_RESIDENT=0 timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2b_bench_c5_stream.json 2> gpurun_out/r2b_bench_c5_stream.err; echo "c5s rc=$?"
timeout 300 python tools/profile_e2e.py > gpurun_out/r2b_profile_e2e.json 2> gpurun_out/r2b_profile_e2e.err; echo "prof rc=$?"
for f in c3_multi c3_g1 c3_multi_pair c4_pair; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b_bench_$f.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"],3), d["stage_ms"], d["check"], d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/r2b_bench_$f.err; done
for f in c5_res c5_stream; do echo "== $f"; head -c 400 gpurun_out/r2b_bench_$f.json; echo; tail -3 gpurun_out/r2b_bench_$f.err; done
cat gpurun_out/r2b_profile_e2e.json; tail -3 gpurun_out/r2b_profile_e2e.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2_ncu_scan_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/r2_ncu_scan_full.log 2>&1; echo "ncu scan rc=$?"
tail -3 gpurun_out/r2_ncu_scan_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2_ncu_scan_c3 python bench.py --workload c3 --chunks 200000 --steps 1 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/r2_ncu_scan_c3.log 2>&1; echo "ncu c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
