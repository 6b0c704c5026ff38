#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python profiles/step_timeline.py > gpurun_out/step_timeline_r3o.json 2> gpurun_out/step_timeline_r3o.err; tail -3 gpurun_out/step_timeline_r3o.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/step_timeline_r3o.json'))
for k in ('span_ms','launches','streams','busy_ms_per_stream','ms_with_0_1_2plus_kernels_running','ms_with_a_gemm_running','ms_with_two_gemms_running','longest_gaps_ms'):
    print(k, d[k])
for k,v in d['per_class'].items(): print(f"{k:24s} {v}")
PY
