#!/bin/bash
# host side of a step: one stage_inputs launch instead of seven copies, eps borrowed instead of cloned -- tests, A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r3m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3m_tests.log
tail -6 gpurun_out/r3m_tests.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3m_$name.json 2> gpurun_out/bench_r3m_$name.err; python - gpurun_out/bench_r3m_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run staged X=1
run copies VDM_STAGE_INPUTS=0
run staged_again X=1
run copies_again VDM_STAGE_INPUTS=0
