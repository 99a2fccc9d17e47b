#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_tests_all.log
tail -6 gpurun_out/r2_tests_all.log
cat gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl
timeout 300 python scripts/quick_xl.py > gpurun_out/r2_quick_main.log 2>&1
tail -5 gpurun_out/r2_quick_main.log
