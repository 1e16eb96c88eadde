set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
date
timeout 300 python -m pytest tests/test_gpu_search.py -m gpu -q -k "tensor_map" --timeout 200 -p no:cacheprovider > gpurun_out/r2u_pytest_tma.log 2>&1; echo "pytest tma rc=$?"; tail -5 gpurun_out/r2u_pytest_tma.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2u_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2u_smoke.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2u_ncu_fp16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check --storage fp16 > gpurun_out/r2u_ncu_fp16.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2u_ncu_fp16.ncu-rep --page raw --csv > gpurun_out/r2u_ncu_fp16_raw.csv 2>/dev/null
ncu -i gpurun_out/r2u_ncu_fp16.ncu-rep --page source --csv > gpurun_out/r2u_ncu_fp16_source.csv 2>/dev/null
rm -f gpurun_out/r2u_ncu_fp16.ncu-rep
timeout 500 ncu --set full --clock-control none -k regex:scan_tcgen05 -s 7 -c 1 -f -o gpurun_out/r2u_ncu_c3_fp16 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-check --storage fp16 > gpurun_out/r2u_ncu_c3_fp16.log 2>&1; echo "ncu c3 rc=$?"
ncu -i gpurun_out/r2u_ncu_c3_fp16.ncu-rep --page raw --csv > gpurun_out/r2u_ncu_c3_fp16_raw.csv 2>/dev/null
rm -f gpurun_out/r2u_ncu_c3_fp16.ncu-rep
date
