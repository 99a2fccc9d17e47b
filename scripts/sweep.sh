#!/bin/bash
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for cfg in "CAR_PDL=0 CAR_L2PF=0" "CAR_PDL=1 CAR_L2PF=0" "CAR_PDL=1 CAR_L2PF=1" "CAR_PDL=1 CAR_L2PF=1 CAR_NSPLIT=1" "CAR_PDL=1 CAR_L2PF=1 CAR_NSPLIT=3"; do
  echo "== $cfg"
  env $cfg N=1024 timeout 120 python scripts/quick_xl.py 2>&1 | grep -E "iter 2|prefill"
  env $cfg N=256 timeout 120 python scripts/quick_xl.py 2>&1 | grep -E "iter 2"
done
for cfg in "CAR_SKIP=31" "CAR_SKIP=30" "CAR_SKIP=29" "CAR_SKIP=27" "CAR_SKIP=23" "CAR_SKIP=15"; do
  echo "== $cfg"
  env CAR_PDL=0 CAR_L2PF=0 $cfg N=256 timeout 120 python scripts/quick_xl.py 2>&1 | grep -E "iter 2"
done
