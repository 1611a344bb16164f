#!/bin/bash
run() {
  python bench.py --steps 4 --warmup 2 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    k = d['kernel_ms_per_step']
    print('$*', json.dumps(dict(tok_s=round(d['value']), ms=round(d['ms_per_step'],2), k={a.replace('o3v_lmhead_',''): round(b,2) for a,b in k.items() if b > 1}, sm=d['clocks']['sm_mhz'])))
"
}
run --fuse-dlogits 0
run --fuse-dlogits 1
run --fuse-dlogits 1 --tunable bwd_prefetch=4
run --fuse-dlogits 1 --tunable bwd_prefetch=8
run --fuse-dlogits 1 --tunable bwd_prefetch=16
run --fuse-dlogits 0 --tunable bwd_prefetch=8
run --fuse-dlogits 0
