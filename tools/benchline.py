"""Print the essentials of bench.py JSON lines: python tools/benchline.py file..."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline", {})
        print("%-40s %-6s value %.3f G/s  %.1f us/step  kern %s  frac %.3f  e2e %.3f G/s  rate %.2f Hz  syn_ev %.0f M/s" % (
            path.split("/")[-1], d["config"].get("method"), d["value"] / 1e9, d.get("us_per_time_step", 0),
            {k: round(v, 1) for k, v in r.get("us_per_launch", {}).items()}, r.get("frac", 0),
            d.get("e2e", {}).get("value", 0) / 1e9, d.get("mean_rate_hz", 0), d.get("syn_events_per_s", 0) / 1e6))
