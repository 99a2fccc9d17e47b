#!/bin/bash
# same-box A/B of decode-kernel code-generation variants (dev): every library under controlar_b200/lib/ab/ + the tree's own build
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for round in 1 2; do
  for L in controlar_b200/lib/ab/*.so ""; do
    echo "variant ${L:-HEAD} round $round" | tee -a gpurun_out/ab.log
    if [ -n "$L" ]; then L=$PWD/$L; fi
    ITERS=2 CAR_LIB=$L timeout 300 python scripts/quick_xl.py 2>&1 | grep -E "iter 1" | tee -a gpurun_out/ab.log
  done
done
