#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "output_head or rowbias" > gpurun_out/r4e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r4e_tests.log
tail -2 gpurun_out/r4e_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r4e_model_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r4e_model_tests.log
tail -2 gpurun_out/r4e_model_tests.log
timeout 300 python profiles/head_conv_time.py > gpurun_out/head_conv_time_r4e.log 2>&1; cat gpurun_out/head_conv_time_r4e.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r4e_$name.json 2> gpurun_out/bench_r4e_$name.err; python - gpurun_out/bench_r4e_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run new X=1
run oldhead VDM_LIB=$PWD/profiles/_diag/libvdm_oldhead.so
run new_again X=1
run oldhead_again VDM_LIB=$PWD/profiles/_diag/libvdm_oldhead.so
