#!/bin/bash
# ncu --set full of the same slice launch (the 100th) for P = 1 and P = 2 (fast build)
export NSB200_LIB=${NSB200_LIB:-$PWD/jaxns_b200/exp/libnsb200_fast.so}
for sp in 1 2; do
  NSB200_SPEC=$sp NSB200_GEN_SMS=120 timeout 280 ncu --set full --clock-control none --import-source on -k regex:k_slice_chains -s 100 -c 1 \
     -o gpurun_out/slice_spec$sp -f python profiles/quick_cfg2.py 1 > gpurun_out/ncu_spec$sp.log 2>&1
  python profiles/ncu_summary.py gpurun_out/slice_spec$sp.ncu-rep > gpurun_out/slice_spec$sp.txt 2>&1
  cat gpurun_out/slice_spec$sp.txt
done
