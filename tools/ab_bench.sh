#!/bin/bash
# A/B of kernel build variants on one GPU box: tools/ab_bench.sh "<nvcc extra flags>" ...
# Each variant is built in place on the box (nvcc is there) and benched with the default workload.
mkdir -p gpurun_out
i=0
for flags in "$@"; do
  out=/tmp/libgelato_ab_$i.so
  python -c "from gelato_b200 import engine; engine.build_library(out='$out', extra='$flags'.split())" || continue
  echo "== variant $i: $flags"
  GELATO_B200_LIB=$out timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | tee gpurun_out/ab_$i.json | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print(b['kernels'], 'step', b['ms_per_step'], 'e2e', b['e2e']['ms_per_step'], b['e2e'].get('full_copy',{}).get('ms_per_step'))"
  i=$((i+1))
done
