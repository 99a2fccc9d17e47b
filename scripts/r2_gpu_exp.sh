#!/bin/bash
# same-box sweep of the pre-poll mode / back-off of the persistent decode kernel (CAR_EXP bits 0-1 = mode, bits 8-11 = sleep / 20 ns)
mkdir -p gpurun_out; rm -f gpurun_out/exp.log
for e in 0 2 514 2562 256 768 1; do
  echo "CAR_EXP=$e" | tee -a gpurun_out/exp.log
  ITERS=3 CAR_EXP=$e timeout 300 python scripts/quick_xl.py 2>&1 | grep -E "iter [12]" | tee -a gpurun_out/exp.log
done
