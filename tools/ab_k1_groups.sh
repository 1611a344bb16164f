#!/bin/bash
# K1 work decomposition at vocab-sliced shapes (one rank's share of c3 at 4 GPUs, 149 vs 148 vocab tiles), 1 GPU:
# balanced groups + rotation (default) vs no rotation vs 8 groups; then DRAM bytes of the GEMM launches (one ncu pass).
out=gpurun_out/r2_ab_k1_groups.jsonl
: > $out
for spec in "c3s4a|fwd_rotate=1" "c3s4a|fwd_rotate=0" "c3s4a|fwd_groups=8" "c3s4a|fwd_groups=8 hint_fwd_store=1" "c3s4b|fwd_rotate=1" "c3s4b|fwd_groups=8" "c2|fwd_rotate=1" "c2|fwd_rotate=0"; do
for once in 1; do
  cfg=${spec%%|*}; tun=${spec##*|}
  args=""
  for kv in $tun; do args="$args --tunable $kv"; done
  python bench.py --config $cfg --steps 4 --warmup 2 --no-cpu-baseline --no-parity $args 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    k = d['kernel_ms_per_step']
    print(json.dumps(dict(cfg='$cfg', tun='$tun', tok_s=round(d['value']), ms=round(d['ms_per_step'],2), k1=round(k.get('o3v_lmhead_fwd+store',0),2), k2a=round(k.get('o3v_lmhead_bwd_dhidden_exp',0),2), k2b=round(k.get('o3v_lmhead_bwd_dweight_exp',0),2), sm=d['clocks']['sm_mhz'])))
" | tee -a $out
done
done
for tun in "fwd_rotate=1" "fwd_rotate=0" "fwd_groups=8"; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lmhead_gemm_kernel -c 5 --csv --log-file gpurun_out/r2_ncu_k1_groups_${tun}.csv python bench.py --config c3s4a --steps 1 --warmup 1 --no-cpu-baseline --no-parity --tunable $tun > /dev/null 2>&1
done
