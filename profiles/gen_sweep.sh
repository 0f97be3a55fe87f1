#!/bin/bash
# A/B of the chain-stream generator schedule on config 2 (device-timed, profiles/quick_cfg2.py)
./profiles/microbench/lat
echo "GEN_MODE=3 (default)"; timeout 120 python profiles/quick_cfg2.py 4 2>&1 | tail -1
echo "GEN_MODE=3 GEN_SMS=120"; NSB200_GEN_SMS=120 timeout 120 python profiles/quick_cfg2.py 4 2>&1 | tail -1
for cfg in "148 128" "148 64" "296 64" "296 128" "148 256" "444 32" "148 32"; do set -- $cfg
  echo "GEN_MODE=4 GEN_SMS=$1 GEN_TPB=$2"; NSB200_GEN_MODE=4 NSB200_GEN_SMS=$1 NSB200_GEN_TPB=$2 timeout 120 python profiles/quick_cfg2.py 4 2>&1 | tail -1
done
