#!/bin/bash
# dev tool: decode-loop timing across the launch strategies (persistent kernel with/without L2 prefetch, graph chain)
for cfg in "CAR_MEGA=1 CAR_MEGA_PF=0" "CAR_MEGA=1 CAR_MEGA_PF=1" "CAR_MEGA=0"; do
  echo "== $cfg"
  env $cfg N=1024 timeout 120 python scripts/quick_xl.py 2>&1 | grep -E "iter 2|prefill|Error|error" | head -5
  env $cfg N=256 timeout 120 python scripts/quick_xl.py 2>&1 | grep -E "iter 2" 
done
