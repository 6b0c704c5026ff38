#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python profiles/ncu_attn_sampler_probe.py --once > gpurun_out/r3u_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'attn_temporal_fused|attn_spatial_sm100' -o gpurun_out/ncu_attn_r3u -f python profiles/ncu_attn_sampler_probe.py --once > gpurun_out/r3u_ncu.log 2>&1
tail -2 gpurun_out/r3u_ncu.log; ls -la gpurun_out/*.ncu-rep
