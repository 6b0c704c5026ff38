#!/bin/bash
# GEMM kernels capped at 128 / 104 registers per thread (__maxnreg__) so that two / three gn_apply blocks of the other
# micro-batch fit beside a persistent GEMM CTA: step A/B through VDM_LIB
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/bench_r3l_$name.json 2> gpurun_out/bench_r3l_$name.err; python - gpurun_out/bench_r3l_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d['clocks']; print(f"{sys.argv[2]:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {c['sm_mhz']} power {c.get('power_w')}")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run base X=1
run r128 VDM_LIB=$PWD/profiles/_diag/libvdm_r128.so
run r104 VDM_LIB=$PWD/profiles/_diag/libvdm_r104.so
run base_again X=1
run r128_again VDM_LIB=$PWD/profiles/_diag/libvdm_r128.so
run r128_mb1 VDM_LIB=$PWD/profiles/_diag/libvdm_r128.so VDM_MICRO_BATCHES=1
