#!/bin/bash
# usage: tools/scale_run.sh "<cfg>:<N> <cfg>:<N> ..." [extra bench args]   (under gpurun --gpus >= max N)
# One bench.py line per (config, N) into gpurun_out/r2_scale_<cfg>_n<N>.json; N = 1 runs in-process, N > 1 via torchrun.
specs="$1"; shift
port=29700
for spec in $specs; do
  cfg=${spec%%:*}; n=${spec##*:}
  out=gpurun_out/r2_scale_${cfg}_n${n}.json
  port=$((port + 1))
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --config $cfg --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out 2> ${out%.json}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --config $cfg --steps 5 --warmup 3 "$@" > $out 2> ${out%.json}.err
  fi
  python - "$out" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["kernel_ms_per_step"]
    print(sys.argv[1], "tok/s %.0f  ms %.2f  e2e %.0f  parity %.3f  frac_burst %.3f  sm %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_max_err"], d["frac_of_bf16_peak"]["burst"], d["clocks"]["sm_mhz"]),
        {a.replace("o3v_lmhead_", ""): round(b, 2) for a, b in k.items() if b > 0.3})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
