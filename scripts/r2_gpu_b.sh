#!/bin/bash
# full GPU test suite + bench line (config 2, incl. the reference GPU-eager arm) + ncu capture of a 24-token launch of the decode kernel
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests_all.log
tail -6 gpurun_out/r2_tests_all.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
tail -c 3000 gpurun_out/r2_bench_c2.json; tail -5 gpurun_out/r2_bench_c2.err
N=24 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pk_decode_kernel -s 1 -c 1 -o gpurun_out/prof_pk_r2 python scripts/quick_xl.py > gpurun_out/ncu_pk_r2.log 2>&1
tail -3 gpurun_out/ncu_pk_r2.log
