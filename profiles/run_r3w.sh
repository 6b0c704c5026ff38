#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "linear or gemm" > gpurun_out/r3w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3w_tests.log
tail -4 gpurun_out/r3w_tests.log
timeout 300 python profiles/gemm_fixed_cost.py > gpurun_out/gemm_fixed_cost_r3w.log 2>&1; tail -9 gpurun_out/gemm_fixed_cost_r3w.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3w_$name.json 2> gpurun_out/bench_r3w_$name.err; python - gpurun_out/bench_r3w_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run new X=1
run astat_off VDM_GEMM_ASTAT=0
run new_again X=1
