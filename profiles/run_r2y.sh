#!/bin/bash
# sustained (200-step) A/B runs at the power cap: identity-K-range residual on/off, micro-batches, pipelined norm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r2y_$name.json 2> gpurun_out/bench_r2y_$name.err; python - gpurun_out/bench_r2y_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run base X=1
run no_identity VDM_IDENTITY_RES_MIN_HW=100000000
run identity_all VDM_IDENTITY_RES_MIN_HW=64
run mb1 VDM_MICRO_BATCHES=1
run pipe1 VDM_PIPELINE_NORM=1
run base_again X=1
