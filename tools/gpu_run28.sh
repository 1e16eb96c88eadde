set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --burst-probe > gpurun_out/r2z_bench_burst.json 2> gpurun_out/r2z_bench_burst.err; echo "rc=$?"; tail -2 gpurun_out/r2z_bench_burst.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_burst.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"]["main_scan"], d["burst_probe"], d["clocks"])
PY
timeout 500 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --burst-probe --storage fp16 > gpurun_out/r2z_bench_burst_fp16.json 2> gpurun_out/r2z_bench_burst_fp16.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_burst_fp16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"]["main_scan"], d["burst_probe"], d["clocks"])
PY
