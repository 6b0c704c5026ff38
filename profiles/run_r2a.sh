#!/bin/bash
# round 2, first GPU call: parity suite, baseline bench + breakdown, ncu rows for attention / sampler / gn_apply
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/kb_r2a.json > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -c 600 gpurun_out/bench_r2a.json
python profiles/ncu_attn_sampler_probe.py > gpurun_out/probe_r2a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'attn_|sampler_step|gn_apply' \
    -o gpurun_out/ncu_attn_sampler_r2a -f python profiles/ncu_attn_sampler_probe.py --once > gpurun_out/ncu_r2a.log 2>&1
cat gpurun_out/probe_r2a.log
tail -3 gpurun_out/ncu_r2a.log
