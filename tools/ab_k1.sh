#!/bin/bash
# A/B of K1 tile mode x vocab groups inside the c2 bench step (same box, back to back).  Run under gpurun.
out=gpurun_out/r2_ab_k1.jsonl
: > $out
for rep in 1 2; do
for cfg in "cta_pair_fwd=1 fwd_groups=4" "cta_pair_fwd=2 fwd_groups=4" "cta_pair_fwd=1 fwd_groups=2" "cta_pair_fwd=2 fwd_groups=2" "cta_pair_fwd=2 fwd_groups=1" "cta_pair_fwd=1 fwd_groups=1" "cta_pair_fwd=2 fwd_groups=3"; do
  args=""
  for kv in $cfg; do args="$args --tunable $kv"; done
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline $args 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    k = d['kernel_ms_per_step']
    print(json.dumps(dict(cfg='$cfg', tok_s=round(d['value']), ms=round(d['ms_per_step'],2), k1=round(k.get('o3v_lmhead_fwd+store',0),2), k2a=round(k.get('o3v_lmhead_bwd_dhidden',0),2), k2b=round(k.get('o3v_lmhead_bwd_dweight',0),2), dl=round(k.get('o3v_lmhead_dlogits',0),2), sm=d['clocks']['sm_mhz'], e2e=round(d['e2e']['value']))))
" | tee -a $out
done
done
