import json
import sys
for line in sys.stdin:
    try:
        j = json.loads(line)
    except ValueError:
        continue
    if "particles_per_s" in j:
        print("  ", j["config"], round(j["device_ms"], 2), "ms", "%.3e" % j["particles_per_s"], "launches", j["launches"])
    else:
        print("  ", line.strip()[:160])
