#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "full_size or forward_matches" > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_tests.log
tail -4 gpurun_out/r2h_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu-baseline"
run() { name=$1; shift; env "$@" timeout 600 $B --profile-json gpurun_out/kb_r2h_$name.json > gpurun_out/bench_r2h_$name.json 2> gpurun_out/bench_r2h_$name.err; }
run pid1 VDM_PROJ_IDENTITY=1
run pid0 VDM_PROJ_IDENTITY=0
run pid1b VDM_PROJ_IDENTITY=1
python - <<'PY'
import json
for f in ('pid1', 'pid0', 'pid1b'):
    try:
        d = json.loads(open(f'gpurun_out/bench_r2h_{f}.json').read().strip().splitlines()[-1])
        k = json.load(open(f'gpurun_out/kb_r2h_{f}.json'))
        proj = sum(g['ms_total'] for g in k['gemm_shapes'] if g['kernel'] == 'gemm_tc_linear' and ('N=384 K=384 ' in g['shape'] or 'N=384 K=768' in g['shape'] or 'N=512 K=512 HxW=8x8' in g['shape'] or 'N=512 K=1024' in g['shape']))
        print(f'{f:8s} ms/step {d["ms_per_step"]:.3f}  e2e {d["e2e"]["ms_per_step"]:.3f}  frac {d["roofline"]["frac"]:.4f}  proj_out launches total {proj:.3f} ms')
    except Exception as e:
        print(f, 'failed', e, open(f'gpurun_out/bench_r2h_{f}.err').read()[-400:])
PY
