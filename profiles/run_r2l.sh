#!/bin/bash
# programmatic dependent launch with the LATE trigger (build -DVDM_PDL_LATE): GEMMs release their dependents when a CTA's
# producer has issued its last load; every other kernel at exit
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r2l_$name.json 2> gpurun_out/bench_r2l_$name.err; python - gpurun_out/bench_r2l_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {d['clocks']['sm_mhz']}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
L=$PWD/video_diffusion_b200/libvdm_pdl.so
run base X=1
run late_pdl_mb2 VDM_LIB=$L VDM_PDL=1
run late_pdl_mb1 VDM_LIB=$L VDM_PDL=1 VDM_MICRO_BATCHES=1
run late_lib_nopdl VDM_LIB=$L
run base_again X=1
VDM_LIB=$L VDM_PDL=1 timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -2
