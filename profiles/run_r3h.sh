#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "temporal or rpe" > gpurun_out/r3h_tests_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r3h_tests_kernels.log
tail -4 gpurun_out/r3h_tests_kernels.log
timeout 300 python profiles/temporal_fused_microbench.py > gpurun_out/temporal_fused_microbench_r3h.json 2> gpurun_out/temporal_fused_microbench_r3h.err
grep -v "^  \"M\"\|max_abs\|hbm_floor" gpurun_out/temporal_fused_microbench_r3h.json; tail -5 gpurun_out/temporal_fused_microbench_r3h.err
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3h_$name.json 2> gpurun_out/bench_r3h_$name.err; python - gpurun_out/bench_r3h_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run default X=1
run pt8 VDM_TEMPORAL_PT=8
run default_again X=1
