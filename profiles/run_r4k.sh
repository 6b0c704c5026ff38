#!/bin/bash
# first micro-batch stream at high priority vs both at default priority
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_r4k_$name.json 2> gpurun_out/bench_r4k_$name.err
python - gpurun_out/bench_r4k_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']
    print(f"{sys.argv[2]:12s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f} clk {c['sm_mhz']}")
except Exception as e: print(sys.argv[2],'failed',e)
PY
}
run default X=1
run mbprio VDM_MB_PRIORITY=1
run default2 X=1
run mbprio2 VDM_MB_PRIORITY=1
