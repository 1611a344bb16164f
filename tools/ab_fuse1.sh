#!/bin/bash
for f in 1 0; do
  python bench.py --steps 4 --warmup 2 --no-cpu-baseline --fuse-dlogits $f 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    k = d['kernel_ms_per_step']
    print(json.dumps(dict(fuse=$f, tok_s=round(d['value']), ms=round(d['ms_per_step'],2), k={a: round(b,2) for a,b in k.items()}, sm=d['clocks']['sm_mhz'])))
"
done
