#!/bin/bash
# relaxed remote arrive when handing a TMEM stage back (pair kernels) + TMA-store epilogue waits for its staging area
# under the first chunk's loads: GEMM tests, linears microbench, step A/B against the previous build is by history
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv or linear or gemm or upsample" > gpurun_out/r3q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3q_tests.log
tail -4 gpurun_out/r3q_tests.log
timeout 300 python profiles/gemm_fixed_cost.py > gpurun_out/gemm_fixed_cost_r3q.log 2>&1; tail -9 gpurun_out/gemm_fixed_cost_r3q.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3q_$name.json 2> gpurun_out/bench_r3q_$name.err; python - gpurun_out/bench_r3q_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run new X=1
run new_again X=1
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r3q_tests_model.log 2>&1; echo "rc=$?" >> gpurun_out/r3q_tests_model.log; tail -3 gpurun_out/r3q_tests_model.log
