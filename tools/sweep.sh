#!/bin/bash
# Runs the smoke parity check + a 250k-query bench for every tuning build in blutils_b200/variants/.
cd "$(dirname "$0")/.."
for so in blutils_b200/variants/*.so; do
  export BLU_CONSENSUS_LIB=$PWD/$so
  ok=$(timeout 120 python __graft_entry__.py smoke 2>&1 | grep -c "smoke ok")
  line=$(timeout 300 python bench.py --queries 250000 --steps 3 --warmup 3 --no-cpu-baseline --no-file-arm 2>/dev/null | tail -1)
  python - "$so" "$ok" "$line" <<'PY'
import json,sys
so,ok,line=sys.argv[1:4]
try:
    d=json.loads(line); print(f"{so.split('/')[-1]:36s} smoke={ok} verified={d['verified']['ok']} tile_ms={d['roofline']['ms_per_launch']:.3f} frac={d['roofline']['frac']:.4f} step_ms={d['ms_per_step']:.3f} e2e={d['e2e']['value']/1e6:.2f}M")
except Exception as e:
    print(so, 'FAILED', ok, line[:200])
PY
done
