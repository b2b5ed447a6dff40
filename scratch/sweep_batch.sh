#!/bin/bash
# experiment: pairs per lock-step LightGlue batch inside the library (bench line without the e2e leg)
for b in "$@"; do
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --lg-batch $b 2>/dev/null > /tmp/sweep_$b.json
  python - "$b" <<'PY'
import json, sys
b = sys.argv[1]
d = json.load(open(f"/tmp/sweep_{b}.json"))
fam = {k.split(" ")[0]: round(v["ms_per_step"], 1) for k, v in d["extra"]["kernel_family_ms_per_step"].items()}
print("batch", b, "pairs/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "flash", round(d["roofline"]["kernel_ms_per_step"], 1), fam, flush=True)
PY
done
