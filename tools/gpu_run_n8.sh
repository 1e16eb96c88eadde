set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2n_bench_c4_n$N.json 2> gpurun_out/r2n_bench_c4_n$N.err; echo "c4 n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n_bench_c4_n$N.json").read().strip().splitlines()[-1]); print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["check"], d["multi_gpu_stage_ms"], d["clocks"])
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/r2n_bench_c4_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus $N --workload c5 > gpurun_out/r2n_bench_c5_n$N.json 2> gpurun_out/r2n_bench_c5_n$N.err; echo "c5 n$N rc=$?"
head -c 400 gpurun_out/r2n_bench_c5_n$N.json; tail -3 gpurun_out/r2n_bench_c5_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2n_bench_ref_n$N.json 2> gpurun_out/r2n_bench_ref_n$N.err; echo "ref n$N rc=$?"; head -c 300 gpurun_out/r2n_bench_ref_n$N.json
