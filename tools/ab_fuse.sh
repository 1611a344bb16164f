#!/bin/bash
# A/B inside the c2 bench step: softmax backward fused into the backward GEMMs vs the separate dlogits pass.
out=gpurun_out/r2_ab_fuse.jsonl
: > $out
for rep in 1 2 3; do
for f in 1 0; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fuse-dlogits $f 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    k = d['kernel_ms_per_step']
    print(json.dumps(dict(fuse=$f, tok_s=round(d['value']), ms=round(d['ms_per_step'],2), loss=d['config']['loss'], k={a: round(b,2) for a,b in k.items()}, sm=d['clocks']['sm_mhz'], e2e=round(d['e2e']['value']))))
" | tee -a $out
done
done
