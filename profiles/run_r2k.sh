#!/bin/bash
# co-residency experiments: gn_apply register cap / block size beside the persistent GEMMs of the other micro-batch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r2k_$name.json 2> gpurun_out/bench_r2k_$name.err; python - gpurun_out/bench_r2k_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {d['clocks']['sm_mhz']}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run mb2 X=1
run mb1 VDM_MICRO_BATCHES=1
run mb2_gn4 VDM_LIB=$PWD/video_diffusion_b200/libvdm_gn4.so
run mb2_gn4_t128 VDM_LIB=$PWD/video_diffusion_b200/libvdm_gn4.so VDM_GN_THREADS=128
run mb2_t128 VDM_GN_THREADS=128
run mb2_again X=1
