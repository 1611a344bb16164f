#!/bin/bash
# usage: tools/ab_dh_collective.sh <N> "<cfg> ..." ["<mode> ..."]   (under gpurun --gpus N)
# A/B of the dHidden collective of the vocab-parallel step at N ranks: reduce-scatter fused into the K2a epilogue
# (NVLink stores to the token owners) vs the one-shot P2P all-reduce beside the dW GEMM, same box, interleaved.
n=$1; cfgs=$2
out=gpurun_out/r2_ab_dh_collective_n${n}.jsonl
: > $out
port=29800
for cfg in $cfgs; do
for mode in ${3:-reduce_scatter reduce_scatter_fused all_reduce}; do
  port=$((port + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --config $cfg --steps 5 --warmup 3 --no-parity --dh-collective $mode 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    r = d.get('kernel_ms_per_step_ranks', {})
    print(json.dumps(dict(cfg='$cfg', mode='$mode', n=$n, tok_s=round(d['value']), ms=round(d['ms_per_step'],2), e2e=round(d['e2e']['value']),
          sm=d['clocks']['sm_mhz'], ranks={k.replace('o3v_lmhead_',''): v for k, v in r.items() if max(v) > 0.3})))
" | tee -a $out
done
done
