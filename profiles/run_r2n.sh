#!/bin/bash
# tcgen05 spatial attention + linears on the transposed-role kernel: full GPU tests, then A/B benches with breakdowns
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_tests.log
tail -4 gpurun_out/r2n_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" $B --profile-json gpurun_out/kb_r2n_$name.json > gpurun_out/bench_r2n_$name.json 2> gpurun_out/bench_r2n_$name.err; python - gpurun_out/bench_r2n_$name.json $name gpurun_out/kb_r2n_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(f"{sys.argv[2]:16s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  frac {d['roofline']['frac']:.4f}  clk {d['clocks']['sm_mhz']}")
    k=json.load(open(sys.argv[3]))
    c=k['per_kernel_class_per_step']
    print('   ', '  '.join(f"{n} {v['ms']:.3f}" for n,v in sorted(c.items(), key=lambda kv:-kv[1]['ms'])[:7]), ' eager sum', round(k['ms_per_step_eager_sum'],3))
    for s in k['gemm_shapes']:
        if s['kernel']=='gemm_tc_linear' and ('N=1152' in s['shape'] or 'N=1536' in s['shape'] or ('res=1 stats=1' in s['shape'])):
            print(f"        {s['shape']:60s} x{s['launches']:3d} {1e3*s['ms_total']/s['launches']:7.1f} us {s['tflops']:6.0f} TF")
except Exception as e: print(sys.argv[2], 'failed', e)
PY
}
run new X=1
run lint0 VDM_GEMM_LINT=0
run attn0 VDM_ATTN_SM100=0
run new_again X=1
