#!/bin/bash
# session-2 GPU call 2: helper-first attention + deferred stream issue: decode parity tests, timing, timeline, ncu source-level capture
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl gpurun_out/resize.jsonl
timeout 900 python -m pytest tests/test_ar_gpu.py tests/test_zz_xl_parity_gpu.py tests/test_zz_resize_gpu.py tests/test_zz_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/f_tests.log
tail -6 gpurun_out/f_tests.log; cat gpurun_out/resize.jsonl
timeout 300 python scripts/quick_xl.py 2>&1 | tail -5 | tee gpurun_out/f_quick_main.log
N=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pk_decode_kernel -s 1 -c 1 -o gpurun_out/prof_pk_f python scripts/quick_xl.py > gpurun_out/ncu_pk_f.log 2>&1
tail -2 gpurun_out/ncu_pk_f.log; ls -la gpurun_out/prof_pk_f.ncu-rep
cp controlar_b200/lib/libcontrolar_b200.so /tmp/lib_keep.so; cp controlar_b200/lib/libcontrolar_b200.so.srchash /tmp/lib_keep.hash
CAR_PK_TRACE=1 python -m controlar_b200.build --force > /dev/null 2>&1
for s in 100 900; do
  CAR_PK_TRACE=1 CAR_DBG=$s timeout 300 python scripts/quick_xl.py 2>&1 | grep "^\[pk" | tail -48 > gpurun_out/f_trace_step$s.log
done
cp /tmp/lib_keep.so controlar_b200/lib/libcontrolar_b200.so; cp /tmp/lib_keep.hash controlar_b200/lib/libcontrolar_b200.so.srchash
head -52 gpurun_out/f_trace_step100.log
