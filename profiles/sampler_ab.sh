#!/bin/bash
# Does the NVML clock sampler perturb the timed region?  bench.py with sampling periods of 1 ms, 20 ms and 1 s.
for p in 1 20 1000; do
  BENCH_CLOCK_PERIOD_MS=$p python bench.py --no-extras --no-cpu-baseline 2>/dev/null | grep '^{' > /tmp/line.json
  python - "$p" <<'PY'
import json, sys
d = json.load(open('/tmp/line.json'))
print("sampler period ms", sys.argv[1], "value", round(d["value"], 1), "ms_per_step", round(d["ms_per_step"], 5), "samples", d["clocks"]["samples"])
PY
done
