#!/bin/bash
# gn_apply: whole-block L2 prefetch (VDM_GN_PF_ALL) -- does the pass make more progress beside the other micro-batch's GEMM?
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "groupnorm" > gpurun_out/r3n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3n_tests.log
VDM_GN_PF_ALL=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "groupnorm" >> gpurun_out/r3n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3n_tests.log
tail -5 gpurun_out/r3n_tests.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3n_$name.json 2> gpurun_out/bench_r3n_$name.err; python - gpurun_out/bench_r3n_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run base X=1
run pf_all VDM_GN_PF_ALL=1
run base_again X=1
run pf_all_again VDM_GN_PF_ALL=1
run pf_all_mb1 VDM_GN_PF_ALL=1 VDM_MICRO_BATCHES=1
