set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
rm -f gpurun_out/scale_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c4shard.json 2> gpurun_out/r2_bench_c4shard.err; echo "c4 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --data clustered --no-cpu-baseline > gpurun_out/r2_bench_c4shard_clustered.json 2> gpurun_out/r2_bench_c4shard_clustered.err; echo "c4cl rc=$?"
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "c2 rc=$?"
timeout 400 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "c3 rc=$?"
timeout 300 python bench.py --workload pool --steps 10 --warmup 3 > gpurun_out/r2_bench_pool.json 2> gpurun_out/r2_bench_pool.err; echo "pool rc=$?"
timeout 400 python bench.py --workload c5 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; echo "c5 rc=$?"
for f in c4shard c4shard_clustered c2 c3 pool c5; do echo "== $f"; tail -c 1500 gpurun_out/r2_bench_$f.json; tail -3 gpurun_out/r2_bench_$f.err; done
