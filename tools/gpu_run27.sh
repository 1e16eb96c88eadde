set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
date
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/r2y_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2y_pytest_dist.log; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_bench_ref_n$N.json 2> gpurun_out/r2y_bench_ref_n$N.err; echo "ref n$N rc=$?"; cut -c1-200 gpurun_out/r2y_bench_ref_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench_c4_n$N.json 2> gpurun_out/r2y_bench_c4_n$N.err; echo "c4 n$N rc=$?"; tail -3 gpurun_out/r2y_bench_c4_n$N.err
if [ "$N" = "2" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --workload c5 > gpurun_out/r2y_bench_c5_n$N.json 2> gpurun_out/r2y_bench_c5_n$N.err; echo "c5 n$N rc=$?"; tail -3 gpurun_out/r2y_bench_c5_n$N.err; fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2y_bench_*_n$N.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f.split("r2y_bench_")[1], round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d.get("e2e",{}).get("ms_per_step",0) or 0,3), r.get("bound"), round(r.get("frac",0) or 0,3), d.get("check"), d.get("multi_gpu_stage_ms"), d.get("clocks",{}).get("sm_mhz"))
    except Exception as e: print(f, "ERR", e)
PY
date
