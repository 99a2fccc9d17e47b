#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests_all.log
tail -8 gpurun_out/r2_tests_all.log
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -6 | tee gpurun_out/r2_bench_gemm.log
timeout 300 python scripts/quick_xl.py 2>&1 | tail -5 | tee gpurun_out/r2_quick_main.log
