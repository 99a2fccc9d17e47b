#!/bin/bash
# round-2 measurement call: full GPU suite, smoke, bench lines of configs 2 / 3 / 4 / 4t, launch list, ncu captures
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl gpurun_out/vision512.jsonl gpurun_out/resize.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/final_tests.log
tail -5 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err; tail -c 600 gpurun_out/final_bench_c2.json; tail -3 gpurun_out/final_bench_c2.err
for c in 3 4 4t; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/final_bench_c$c.json 2> gpurun_out/final_bench_c$c.err; tail -c 300 gpurun_out/final_bench_c$c.json; done
timeout 300 python scripts/vision_once.py 2>&1 | tail -3 | tee gpurun_out/final_vision.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --no-gpu-eager --no-cpu-baseline > gpurun_out/final_launches_bench.log 2>&1
gzip -f gpurun_out/final_launches.csv
N=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pk_decode_kernel -s 1 -c 1 -o gpurun_out/final_prof_pk python scripts/quick_xl.py > gpurun_out/final_ncu_pk.log 2>&1
B=2 timeout 900 ncu --set full --clock-control none -k regex:"dense_gemm_kernel|gemm_tc5_kernel" -s 200 -c 12 -o gpurun_out/final_prof_dense python scripts/vision_once.py > gpurun_out/final_ncu_dense.log 2>&1
ls -la gpurun_out/*.ncu-rep
