#!/bin/bash
# session-2 GPU call 1: new tests first (train forward, resize, vision 512), then the full suite, a decode timeline, quick timing
mkdir -p gpurun_out
rm -f gpurun_out/xl_parity.jsonl gpurun_out/small_parity.jsonl gpurun_out/vision512.jsonl
timeout 600 python -m pytest tests/test_zz_train_forward_gpu.py tests/test_zz_resize_gpu.py tests/test_zz_vision512_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/e_new_tests.log
tail -30 gpurun_out/e_new_tests.log
cat gpurun_out/vision512.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_zz_train_forward_gpu.py --deselect tests/test_zz_resize_gpu.py --deselect tests/test_zz_vision512_gpu.py 2>&1 | tail -15 > gpurun_out/e_tests_all.log
tail -6 gpurun_out/e_tests_all.log
timeout 300 python scripts/quick_xl.py 2>&1 | tail -5 | tee gpurun_out/e_quick_main.log
# per-phase timeline (trace build in a scratch copy of the library)
cp controlar_b200/lib/libcontrolar_b200.so /tmp/lib_keep.so; cp controlar_b200/lib/libcontrolar_b200.so.srchash /tmp/lib_keep.hash
CAR_PK_TRACE=1 python -m controlar_b200.build --force > /dev/null 2>&1
for s in 100 900; do
  CAR_PK_TRACE=1 CAR_DBG=$s timeout 300 python scripts/quick_xl.py 2>&1 | grep "^\[pk" | tail -45 > gpurun_out/e_trace_step$s.log
done
cp /tmp/lib_keep.so controlar_b200/lib/libcontrolar_b200.so; cp /tmp/lib_keep.hash controlar_b200/lib/libcontrolar_b200.so.srchash
head -50 gpurun_out/e_trace_step100.log
