#!/bin/bash
# decode-chunk sweep of the NS2d bench (timing experiment)
for c in "$@"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --decode-chunk $c 2>/dev/null > /tmp/b_$c.json
  python - <<PY
import json
d = json.loads(open("/tmp/b_$c.json").read())
print("chunk $c:", d["value"], "traj-steps/s,", d["ms_per_step"], "ms/step,", d["launches_per_step"], "launches")
PY
done
