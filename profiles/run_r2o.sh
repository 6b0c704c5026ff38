#!/bin/bash
# tcgen05 spatial attention: per-launch times (L2 flushed) against the mma.sync kernel, then one ncu --set full capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python profiles/ncu_attn_sampler_probe.py 2>&1 | grep attn_spatial
VDM_ATTN_SM100=0 python profiles/ncu_attn_sampler_probe.py 2>&1 | grep attn_spatial
ncu --set full --clock-control none --import-source on -k regex:'attn_spatial_sm100' -o gpurun_out/ncu_attn_sm100_r2o -f python profiles/ncu_attn_sampler_probe.py --once > gpurun_out/ncu_attn_r2o.log 2>&1
ls -la gpurun_out/
