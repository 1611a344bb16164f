#!/bin/bash
# A/B of the three backward variants inside the c2 bench step (same box, interleaved).
out=gpurun_out/r2_ab_backward.jsonl
: > $out
for rep in 1 2 3; do
for b in exp dlogits; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --backward $b 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    k = d['kernel_ms_per_step']
    print(json.dumps(dict(backward='$b', tok_s=round(d['value']), ms=round(d['ms_per_step'],2), e2e=round(d['e2e']['value']), parity=round(d['parity_max_err'],3), k={a.replace('o3v_lmhead_',''): round(v,2) for a,v in k.items()}, sm=d['clocks']['sm_mhz'], pw=d['clocks'].get('power_w_median'))))
" | tee -a $out
done
done
