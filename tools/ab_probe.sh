#!/bin/bash
# A/B of kernel build variants on one GPU box with the kernel-level probe: tools/ab_probe.sh "<nvcc extra flags>" ...
mkdir -p gpurun_out
i=0
for flags in "$@"; do
  out=/tmp/libgelato_ab_$i.so
  python -c "from gelato_b200 import engine; engine.build_library(out='$out', extra='$flags'.split())" || { i=$((i+1)); continue; }
  echo "== variant $i: $flags" | tee -a gpurun_out/ab_probe.txt
  GELATO_B200_LIB=$out timeout 300 python tools/kernel_probe.py | tee -a gpurun_out/ab_probe.txt
  i=$((i+1))
done
