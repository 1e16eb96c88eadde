cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 70 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --data clustered > gpurun_out/r3b_bench_clustered.json 2> gpurun_out/r3b_bench_clustered.err; echo "clustered rc=$?"
timeout 75 python bench.py --no-cpu-baseline --steps 5 --warmup 3 --filtered > gpurun_out/r3b_bench_filtered.json 2> gpurun_out/r3b_bench_filtered.err; echo "filtered rc=$?"
python - <<'PY'
import json
for n in ("clustered","filtered"):
    try:
        d=json.loads(open(f"gpurun_out/r3b_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["stage_ms"], d["check"]["identical_topk_sets"], d.get("filtered"), d["robustness"])
    except Exception as e: print(n, "ERR", e)
PY
