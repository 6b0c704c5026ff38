#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "a_stationary or tma_store" > gpurun_out/r2i_k.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_k.log
tail -12 gpurun_out/r2i_k.log
if grep -q "rc=0" gpurun_out/r2i_k.log; then
  timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_tests.log
  tail -4 gpurun_out/r2i_tests.log
  B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
  run() { name=$1; shift; env "$@" timeout 600 $B --profile-json gpurun_out/kb_r2i_$name.json > gpurun_out/bench_r2i_$name.json 2> gpurun_out/bench_r2i_$name.err; }
  run as1 VDM_GEMM_ASTAT=1
  run as0 VDM_GEMM_ASTAT=0
  python - <<'PY'
import json
for f in ('as1', 'as0'):
    try:
        d = json.loads(open(f'gpurun_out/bench_r2i_{f}.json').read().strip().splitlines()[-1])
        k = json.load(open(f'gpurun_out/kb_r2i_{f}.json'))
        print(f'{f:8s} ms/step {d["ms_per_step"]:.3f}  e2e {d["e2e"]["ms_per_step"]:.3f}  frac {d["roofline"]["frac"]:.4f}')
        for g in k['gemm_shapes']:
            if g['kernel'] == 'gemm_tc_linear' and g['launches'] >= 5:
                print(f"    {g['shape'][:70]:72s} x{g['launches']:2d} {g['ms_total'] / g['launches'] * 1e3:7.1f} us {g['tflops']:6.0f} TF")
    except Exception as e:
        print(f, 'failed', e, open(f'gpurun_out/bench_r2i_{f}.err').read()[-400:])
PY
fi
