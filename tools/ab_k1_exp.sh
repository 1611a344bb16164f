#!/bin/bash
# A/B of the K1 tile mode (1-CTA tiles vs cta_group::2 pairs) inside the c2 step of the exp-store pipeline
# (same box, interleaved).  Run under gpurun.
out=gpurun_out/r2_ab_k1_pair_expstore.jsonl
: > $out
for rep in 1 2; do
for cfg in "cta_pair_fwd=1" "cta_pair_fwd=2" "cta_pair_fwd=2 fwd_groups=2"; do
  args=""
  for kv in $cfg; do args="$args --tunable $kv"; done
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity $args 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    k = d['kernel_ms_per_step']
    print(json.dumps(dict(cfg='$cfg', tok_s=round(d['value']), ms=round(d['ms_per_step'],2), k1=round(k.get('o3v_lmhead_fwd+store',0),2), k2a=round(k.get('o3v_lmhead_bwd_dhidden_exp',0),2), k2b=round(k.get('o3v_lmhead_bwd_dweight_exp',0),2), sm=d['clocks']['sm_mhz'], power_med=d['clocks'].get('power_w_median'), power_max=d['clocks'].get('power_w_max'), e2e=round(d['e2e']['value']))))
" | tee -a $out
done
done
