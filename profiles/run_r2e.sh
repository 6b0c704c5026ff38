#!/bin/bash
# round 2, call e: TMA-store epilogue for the plain linears
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "tma_store or gemm" > gpurun_out/r2e_k.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_k.log
tail -12 gpurun_out/r2e_k.log
if grep -q "rc=0" gpurun_out/r2e_k.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
  tail -6 gpurun_out/r2e_tests.log
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline --profile-json gpurun_out/kb_r2e.json > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
  VDM_GEMM_TS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline > gpurun_out/bench_r2e_nots.json 2> gpurun_out/bench_r2e_nots.err
  python - <<'PY'
import json
for f in ('bench_r2e', 'bench_r2e_nots'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4), 'gemm ms', round(d['roofline']['gemm']['kernel_ms_per_step'], 3))
    except Exception as e:
        print(f, 'failed', e)
k = json.load(open('gpurun_out/kb_r2e.json'))
for g in k['gemm_shapes']:
    if g['kernel'] == 'gemm_tc_linear':
        print(f"{g['shape'][:72]:74s} x{g['launches']:3d} {g['ms_total'] / g['launches'] * 1e3:7.1f} us {g['tflops']:7.0f} TF")
PY
fi
