#!/bin/bash
# session-2 GPU call 4: sampler tests (standalone kernel), decode parity, timing incl. the L2 run-ahead knob, timeline
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ar_gpu.py tests/test_zz_xl_parity_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/h_tests.log
tail -8 gpurun_out/h_tests.log
timeout 300 python scripts/quick_xl.py 2>&1 | tail -5 | tee gpurun_out/h_quick_main.log
for e in 8208 12304 16400 24592; do echo "CAR_EXP=$e"; CAR_EXP=$e timeout 300 python scripts/quick_xl.py 2>&1 | grep "iter [12]" | tee -a gpurun_out/h_quick_exp.log; done
cp controlar_b200/lib/libcontrolar_b200.so /tmp/lib_keep.so; cp controlar_b200/lib/libcontrolar_b200.so.srchash /tmp/lib_keep.hash
CAR_PK_TRACE=1 python -m controlar_b200.build --force > /dev/null 2>&1
CAR_PK_TRACE=1 CAR_DBG=100 timeout 300 python scripts/quick_xl.py 2>&1 | grep "^\[pk" | tail -55 > gpurun_out/h_trace_step100.log
cp /tmp/lib_keep.so controlar_b200/lib/libcontrolar_b200.so; cp /tmp/lib_keep.hash controlar_b200/lib/libcontrolar_b200.so.srchash
grep -v "warp\]" gpurun_out/h_trace_step100.log
