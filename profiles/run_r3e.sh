#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/temporal_fused_microbench.py > gpurun_out/temporal_fused_microbench_r3e.json 2> gpurun_out/temporal_fused_microbench_r3e.err
cat gpurun_out/temporal_fused_microbench_r3e.json; tail -5 gpurun_out/temporal_fused_microbench_r3e.err
