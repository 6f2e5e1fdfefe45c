#!/bin/bash
# A/B an environment knob on the inference bench: tools/ab_env.sh VAR v1 v2 [repeats]
VAR=$1; A=$2; B=$3; N=${4:-2}
for i in $(seq $N); do for v in $A $B; do
  env $VAR=$v timeout 200 python bench.py --steps 20 --no-train --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/ab.json
  python - "$VAR=$v" <<'PY'
import json, sys
d = json.load(open("/tmp/ab.json"))
print(sys.argv[1], round(d["value"], 1), round(d["roofline"]["frac"], 4), round(d["roofline_aux"]["ms_by_kernel_family"]["conv_gemm"], 2))
PY
done; done
