#!/bin/bash
# round 2, session 3: output head with fused GroupNorm + SiLU (tests, bench A/B); per-role trace of the attention linears
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "output_head" > gpurun_out/r3g_tests_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r3g_tests_kernels.log
tail -6 gpurun_out/r3g_tests_kernels.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r3g_tests_model.log 2>&1; echo "rc=$?" >> gpurun_out/r3g_tests_model.log
tail -4 gpurun_out/r3g_tests_model.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3g_$name.json 2> gpurun_out/bench_r3g_$name.err; python - gpurun_out/bench_r3g_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run head1 X=1
run head0 VDM_FUSE_HEAD=0
run head1_pdl VDM_PDL=1
run head1_again X=1
run head0_again VDM_FUSE_HEAD=0
VDM_LIB=$PWD/profiles/_diag/libvdm_trace.so TRACE_SHAPES=qkv16,proj16,qkv8,proj8 timeout 300 python profiles/gemm_trace.py > gpurun_out/gemm_trace_r3g.log 2>&1; tail -20 gpurun_out/gemm_trace_r3g.log
