#!/bin/bash
# four micro-batches (tensors of a 64x64 stage then fit the L2 between producer and consumer) vs the default two
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r4g_$name.json 2> gpurun_out/bench_r4g_$name.err; python - gpurun_out/bench_r4g_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
}
run mb2 X=1
run mb4 VDM_MICRO_BATCHES=4
run mb4_join256 VDM_MICRO_BATCHES=4 VDM_MB_JOIN_HW=256
run mb4_join1024 VDM_MICRO_BATCHES=4 VDM_MB_JOIN_HW=1024
run mb2_again X=1
