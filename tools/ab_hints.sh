run() { name=$1; shift; timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 3 "$@" > gpurun_out/ab_$name.json 2>gpurun_out/ab_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {k.replace("o3v_lmhead_",""):round(v,1) for k,v in d["kernel_ms_per_step"].items() if v>1})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/ab_$name.err").read()[-500:])
PY
}
run base
run st1 --tunable hint_fwd_store=1
run st1_b2 --tunable hint_fwd_store=1 --tunable hint_fwd_b=2
run st1_a2 --tunable hint_fwd_store=1 --tunable hint_fwd_a=2
run bwd_b2 --tunable hint_bwd_b=2
run bwd_a2 --tunable hint_bwd_a=2
run base2
run st1_2 --tunable hint_fwd_store=1
