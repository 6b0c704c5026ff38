#!/bin/bash
# round 2, call f: fp16 residual stream
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "fp16 or groupnorm or gemm or halo or second_range" > gpurun_out/r2f_k.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_k.log
tail -12 gpurun_out/r2f_k.log
if grep -q "rc=0" gpurun_out/r2f_k.log; then
  timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
  grep -E "max-rel|PSNR|passed|failed|FAILED" gpurun_out/r2f_tests.log | tail -40
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline --profile-json gpurun_out/kb_r2f.json > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
  python - <<'PY'
import json
for f in ('bench_r2f',):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 4), 'gemm ms', round(d['roofline']['gemm']['kernel_ms_per_step'], 3), d['roofline_hbm'].get('gn_apply', {}).get('ms_per_step'))
    except Exception as e:
        print(f, 'failed', e, open(f'gpurun_out/{f}.err').read()[-600:])
k = json.load(open('gpurun_out/kb_r2f.json'))
for n, v in sorted(k['per_kernel_class_per_step'].items(), key=lambda kv: -kv[1]['ms']):
    print(f"{n:24s} {v['ms']:.3f} ms  {v['launches']:.0f}")
PY
fi
