#!/bin/bash
# compute-sanitizer over the kernels on a small problem (run under gpurun): memcheck, racecheck, initcheck.
# The workload is the bit-exact GPU test of the shipped example plus the update-mode and pair tests.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 --error-exitcode 9 \
    python -m pytest -m gpu tests/test_gpu_parity.py -x -q tests/test_initguess.py -k "golden_bitwise or update_mode or pair_evaluation or packed_pair or batched or sliced_update or user_constraint_registry or forward_simulation_kernel" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log) $(grep 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  tail -2 gpurun_out/sanitize_$tool.log | head -1
done
