#!/bin/bash
# device value of bench.py over batch sizes around whole numbers of waves (26 blocks per scenario, 296 resident blocks)
for s in 114 125 128 137 148 159; do
  python bench.py --no-cpu-baseline --sustained-seconds 0.3 --preheat-seconds 0.2 --scenarios $s 2>/dev/null > /tmp/b_$s.json
  python - <<PY
import json
d = json.load(open("/tmp/b_$s.json"))
print("scenarios", $s, "waves", round(26 * $s / 296, 2), "value G", round(d["value"] / 1e9, 3), "ms", round(d["ms_per_step"], 4),
      "sustained G", round(d["sustained"]["value"] / 1e9, 3), "k_jacobian ms", round(d["kernels"]["k_jacobian_ms"], 4),
      "frac", round(d["roofline"]["frac"], 4), "e2e G", round(d["e2e"]["value"] / 1e9, 3))
PY
done
