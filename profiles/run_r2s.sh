#!/bin/bash
# gn_apply with 128-thread blocks and a register cap (more blocks fit beside a persistent GEMM CTA)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r2s_$name.json 2> gpurun_out/bench_r2s_$name.err; python - gpurun_out/bench_r2s_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {d['clocks']['sm_mhz']} gn_apply {d['roofline_hbm']['gn_apply']['ms_per_step']:.3f}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
D=$PWD/video_diffusion_b200
run base X=1
run gn_128_7 VDM_LIB=$D/libvdm_gn_128_7.so
run gn_128_6 VDM_LIB=$D/libvdm_gn_128_6.so
run base_again X=1
run gn_128_7_again VDM_LIB=$D/libvdm_gn_128_7.so
