#!/bin/bash
# same-box A/B of decode-kernel variants (dev): A = round-2 start (92ed9c5), C = new sampler only, B = HEAD (C + RoPE preload)
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for round in 1 2; do
  for v in A C B; do
    if [ $v = B ]; then L=""; else L=$PWD/controlar_b200/lib/ab/lib$v.so; fi
    echo "variant $v round $round" | tee -a gpurun_out/ab.log
    CAR_LIB=$L timeout 300 python scripts/quick_xl.py 2>&1 | grep -E "iter [12]|prefill" | tee -a gpurun_out/ab.log
  done
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv | tee -a gpurun_out/ab.log
