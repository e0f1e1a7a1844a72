#!/bin/bash
# One GPU session: tests, smoke, bench, launch list and ncu captures.  Run under gpurun from the repo root.
#   tools/gpu_check.sh [quick]     quick: tests + bench only (no ncu)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
[ "$1" = "quick" ] && exit 0
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
NCU_BENCH="python bench.py --steps 3 --warmup 3 --solve-scenarios 0 --no-cpu-baseline --sustained-seconds 0.01 --preheat-seconds 0.01"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  $NCU_BENCH > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
for k in k_jacobian k_jacobian_noair k_residuals; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 6 -c 1 -f -o gpurun_out/prof_$k \
    $NCU_BENCH > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out
