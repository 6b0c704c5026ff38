#!/bin/bash
# image-pipelined out_layers GroupNorm-apply (VDM_PIPELINE_NORM): model tests, then A/B benches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r2u_$name.json 2> gpurun_out/bench_r2u_$name.err; python - gpurun_out/bench_r2u_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {d['clocks']['sm_mhz']}")
except Exception as e: print(sys.argv[2], 'failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
run pipe1 X=1
run pipe0 VDM_PIPELINE_NORM=0
run pipe1_mb1 VDM_MICRO_BATCHES=1
run pipe0_mb1 VDM_PIPELINE_NORM=0 VDM_MICRO_BATCHES=1
run pipe1_again X=1
run pipe0_again VDM_PIPELINE_NORM=0
