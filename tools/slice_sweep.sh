#!/bin/bash
# end-to-end leg of bench.py over pipeline slice counts (GPU box)
for s in 2 4 8 16 32; do
  python bench.py --no-cpu-baseline --sustained-seconds 0.05 --preheat-seconds 0.05 --e2e-slices $s 2>/dev/null > /tmp/b_$s.json
  python - <<PY
import json
d = json.load(open("/tmp/b_$s.json"))
print("slices", $s, "e2e ms", round(d["e2e"]["ms_per_step"], 4), "update ms", round(d["e2e"]["coo_update_mode"]["ms_per_step"], 4))
PY
done
