#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/head_conv_ncu.py || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_head -s 2 -c 2 -f -o gpurun_out/ncu_head_r4d python profiles/head_conv_ncu.py > gpurun_out/ncu_head_r4d.log 2>&1
tail -3 gpurun_out/ncu_head_r4d.log
