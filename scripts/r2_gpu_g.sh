#!/bin/bash
# session-2 GPU call 3: rewritten sampler: sampler / decode parity tests, timing, timeline
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ar_gpu.py tests/test_zz_fullsize_gpu.py tests/test_zz_legacy_c2i_gpu.py tests/test_zz_dropin_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/g_tests.log
tail -12 gpurun_out/g_tests.log
timeout 300 python scripts/quick_xl.py 2>&1 | tail -5 | tee gpurun_out/g_quick_main.log
cp controlar_b200/lib/libcontrolar_b200.so /tmp/lib_keep.so; cp controlar_b200/lib/libcontrolar_b200.so.srchash /tmp/lib_keep.hash
CAR_PK_TRACE=1 python -m controlar_b200.build --force > /dev/null 2>&1
for s in 100; do
  CAR_PK_TRACE=1 CAR_DBG=$s timeout 300 python scripts/quick_xl.py 2>&1 | grep "^\[pk" | tail -55 > gpurun_out/g_trace_step$s.log
done
cp /tmp/lib_keep.so controlar_b200/lib/libcontrolar_b200.so; cp /tmp/lib_keep.hash controlar_b200/lib/libcontrolar_b200.so.srchash
grep -v "warp\]" gpurun_out/g_trace_step100.log | grep -E "sampler|head|barrier|step start|w2 end|qkv start"
