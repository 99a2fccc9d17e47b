#!/bin/bash
mkdir -p gpurun_out
CAR_LIB=$PWD/build/libcontrolar_b200_trace.so CAR_DBG=100 N=256 timeout 300 python scripts/quick_xl.py 2>&1 | tail -60 > gpurun_out/r2_trace_pk1.log
CAR_LIB=$PWD/build/libcontrolar_b200_trace.so CAR_DBG=900 timeout 300 python scripts/quick_xl.py 2>&1 | tail -60 > gpurun_out/r2_trace_pk1_n1000.log
for c in 3 4 4t; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err; tail -c 1500 gpurun_out/r2_bench_c$c.json; done
