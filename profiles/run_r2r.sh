#!/bin/bash
# programmatic dependent launch, now also on the halo conv kernels: early trigger everywhere (VDM_PDL=1) vs the mixed
# protocol (-DVDM_PDL_LATE build: persistent GEMMs trigger after their last load, short-block kernels at their top)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r2r_$name.json 2> gpurun_out/bench_r2r_$name.err; python - gpurun_out/bench_r2r_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {d['clocks']['sm_mhz']}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
L=$PWD/video_diffusion_b200/libvdm_pdl.so
run base X=1
run pdl_early VDM_PDL=1
run pdl_mixed VDM_LIB=$L VDM_PDL=1
run pdl_mixed_mb1 VDM_LIB=$L VDM_PDL=1 VDM_MICRO_BATCHES=1
run base_mb1 VDM_MICRO_BATCHES=1
run base_again X=1
run pdl_mixed_again VDM_LIB=$L VDM_PDL=1
VDM_LIB=$L VDM_PDL=1 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -2
