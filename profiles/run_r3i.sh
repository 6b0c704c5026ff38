#!/bin/bash
# round 2, session 3: attention qkv projections reading the fp16 stream copy directly (VDM_QKV_F16) -- tests, bench A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "fp16_operands or only_output or a_stationary or temporal" > gpurun_out/r3i_tests_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r3i_tests_kernels.log
tail -6 gpurun_out/r3i_tests_kernels.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r3i_tests_model.log 2>&1; echo "rc=$?" >> gpurun_out/r3i_tests_model.log
tail -4 gpurun_out/r3i_tests_model.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3i_$name.json 2> gpurun_out/bench_r3i_$name.err; python - gpurun_out/bench_r3i_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run qkv16_on X=1
run qkv16_off VDM_QKV_F16=0
run qkv16_on_again X=1
run qkv16_off_again VDM_QKV_F16=0
