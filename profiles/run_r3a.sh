#!/bin/bash
# round 2, session 3: fused temporal attention kernel -- parity tests, microbench against the three-launch path,
# bench A/B (fused on/off, 8 / 16 pixels per CTA)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "temporal or rpe" > gpurun_out/r3a_tests_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r3a_tests_kernels.log
tail -15 gpurun_out/r3a_tests_kernels.log
timeout 300 python profiles/temporal_fused_microbench.py > gpurun_out/temporal_fused_microbench_r3a.json 2> gpurun_out/temporal_fused_microbench_r3a.err
cat gpurun_out/temporal_fused_microbench_r3a.json; tail -5 gpurun_out/temporal_fused_microbench_r3a.err
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r3a_tests_model.log 2>&1; echo "rc=$?" >> gpurun_out/r3a_tests_model.log
tail -8 gpurun_out/r3a_tests_model.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3a_$name.json 2> gpurun_out/bench_r3a_$name.err; python - gpurun_out/bench_r3a_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run fused_pt8 X=1
run old VDM_FUSED_TEMPORAL=0
run fused_pt16 VDM_TEMPORAL_PT=16
run fused_pt8_again X=1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline --profile-json gpurun_out/kb_r3a.json > gpurun_out/bench_r3a_profile.json 2> gpurun_out/bench_r3a_profile.err
tail -3 gpurun_out/bench_r3a_profile.err
