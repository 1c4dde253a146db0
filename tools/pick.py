"""Print the few numbers of a bench.py JSON line (stdin) that an A/B comparison looks at."""
import json
import sys

lines = [ln for ln in sys.stdin.read().strip().split("\n") if ln.startswith("{")]
if not lines:
    sys.exit("no JSON line")
j = json.loads(lines[-1])
r = j["roofline"]
print("ms/step", round(j["ms_per_step"], 5), "rows_ms", round(r["kernel_ms"], 5), "frac", round(r["frac"], 3),
      "e2e_s", round(j["e2e"]["seconds"], 4), "build_p_ms", round(j.get("build_p", {}).get("ms", 0.0), 4))
