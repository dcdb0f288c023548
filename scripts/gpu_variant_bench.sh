#!/bin/bash
# kernel experiments: the same bench (131072 nodes, no CPU legs) with differently built copies of the library
mkdir -p gpurun_out
for v in "$@"; do
  WTG_LIB=$PWD/wittgenstein_b200/$v timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = [json.loads(l) for l in open(f"gpurun_out/variant_{v}.json") if l.startswith("{")][0]
    r = d["roofline"]
    print(v, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "k_node frac", round(r["frac"], 3), {k: round(x) for k, x in r["kernel_ms"].items()})
except Exception as e:
    print(v, "failed", e, open(f"gpurun_out/variant_{v}.err").read()[-400:])
PY
done
