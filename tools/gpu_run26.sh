set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
date
rm -f gpurun_out/scale_parity.jsonl
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 -p no:cacheprovider > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2x_smoke.log
date
timeout 600 python bench.py --impl reference > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/r2x_bench_reference.json
timeout 600 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err; echo "default rc=$?"; tail -2 gpurun_out/r2x_bench_default.err
date
run() { name=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/r2x_bench_$name.json 2> gpurun_out/r2x_bench_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/r2x_bench_$name.err; }
run c4_fp16 --no-cpu-baseline --steps 20 --warmup 3 --storage fp16
run c3_fp16 --no-cpu-baseline --workload c3 --steps 10 --warmup 3 --storage fp16
run c2 --no-cpu-baseline --workload c2 --steps 20 --warmup 3
run pool --no-cpu-baseline --workload pool
run c5 --workload c5
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2x_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f.split("r2x_bench_")[1], round(d["value"],2), d["unit"][:12], "ms", round(d["ms_per_step"],3), "e2e", round(d.get("e2e",{}).get("ms_per_step",0) or 0,3), r.get("bound"), round(r.get("frac",0),3), d.get("check",{}).get("identical_topk_sets"), d.get("clocks",{}).get("sm_mhz"))
    except Exception as e: print(f, "ERR", e)
PY
date
