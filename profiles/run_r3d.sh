#!/bin/bash
# round 2, session 3: micro-batches that join for the small levels (VDM_MB_JOIN_HW) -- model parity tests, bench A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r3d_tests_model.log 2>&1; echo "rc=$?" >> gpurun_out/r3d_tests_model.log
tail -8 gpurun_out/r3d_tests_model.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3d_$name.json 2> gpurun_out/bench_r3d_$name.err; python - gpurun_out/bench_r3d_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run join256 X=1
run join0 VDM_MB_JOIN_HW=0
run join1024 VDM_MB_JOIN_HW=1024
run join64 VDM_MB_JOIN_HW=64
run join256_again X=1
run join0_again VDM_MB_JOIN_HW=0
run mb1 VDM_MICRO_BATCHES=1
