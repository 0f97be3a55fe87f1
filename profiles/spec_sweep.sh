#!/bin/bash
# A/B of the speculative-proposal width P (NSB200_SPEC) on config 2; NSB200_LIB picks the build under test
export NSB200_LIB=${NSB200_LIB:-$PWD/jaxns_b200/exp/libnsb200_fast.so}
for sp in 1 2 4; do for sms in 120; do
  echo "SPEC=$sp GEN_SMS=$sms"; NSB200_SPEC=$sp NSB200_GEN_SMS=$sms timeout 120 python profiles/quick_cfg2.py 4 2>&1 | tail -1
done; done
