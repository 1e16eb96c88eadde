set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/r2n_pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2n_pytest_dist.log
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2n_bench_c4_n$N.json 2> gpurun_out/r2n_bench_c4_n$N.err; echo "c4 n$N rc=$?"
tail -c 2500 gpurun_out/r2n_bench_c4_n$N.json; tail -3 gpurun_out/r2n_bench_c4_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 --filtered --no-cpu-baseline > gpurun_out/r2n_bench_c4f_n$N.json 2> gpurun_out/r2n_bench_c4f_n$N.err; echo "c4f n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_bench_c4f_n$N.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("filtered"), d["check"], d["multi_gpu_stage_ms"])
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/r2n_bench_c4f_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --workload c5 > gpurun_out/r2n_bench_c5_n$N.json 2> gpurun_out/r2n_bench_c5_n$N.err; echo "c5 n$N rc=$?"
head -c 500 gpurun_out/r2n_bench_c5_n$N.json; tail -3 gpurun_out/r2n_bench_c5_n$N.err
