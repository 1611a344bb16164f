#!/bin/bash
# c2 on one GPU: how the step is cut into token chunks (whole sequences) and how many vocab groups K1 uses.
out=gpurun_out/r2_ab_chunk_plan_c2.jsonl
: > $out
for rep in 1 2; do
for v in "--no-chunk-plan" "--chunk-tokens 32768" "--chunk-tokens 32768 --tunable fwd_groups=4" "" "--tunable fwd_groups=4"; do
  python bench.py --config c2 $v --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    print(json.dumps(dict(args='$v', plan=d['config']['chunk_sequences'], tok_s=round(d['value']), ms=round(d['ms_per_step'],2), e2e=round(d['e2e']['value']), k={k.replace('o3v_lmhead_',''): round(v,2) for k,v in d['kernel_ms_per_step'].items() if v > 0.3}, sm=d['clocks']['sm_mhz'])))
" | tee -a $out
done
done
