set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/scale_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2f_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2f_smoke.log
timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2f_bench_c5_att2.json 2> gpurun_out/r2f_bench_c5_att2.err; echo "c5 rc=$?"; head -c 300 gpurun_out/r2f_bench_c5_att2.json; tail -2 gpurun_out/r2f_bench_c5_att2.err
RL_XENC_ATT2=0 timeout 400 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2f_bench_c5_att1.json 2> gpurun_out/r2f_bench_c5_att1.err; echo "c5 rc=$?"; head -c 300 gpurun_out/r2f_bench_c5_att1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear|attention|add_ln|cls_head|embed_ln" -c 120 --csv --log-file gpurun_out/r2f_launches_xenc.csv python tools/bench_rerank.py --pairs 700 --tokens-per-call 51200 --cpu-pairs 2 > gpurun_out/r2f_launches_xenc.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2f_launches_xenc.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki].split('(')[0].split('::')[-1], float(r[vi].replace(',',''))) for r in rows[1:]]
for s in seq[60:70]: print(s)
PY
