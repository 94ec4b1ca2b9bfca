"""stdin: bench.py output; stdout: ms per step, value and the per-stage kernel times of its JSON line (one short line for A/B logs)"""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        j = json.loads(line)
        print(j.get("ms_per_step"), j.get("value"), " ".join("%s=%.4f" % (k, v["ms"]) for k, v in j.get("stages", {}).items()))
