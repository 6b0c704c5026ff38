#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python profiles/gn_ncu_probe.py > gpurun_out/r3t_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'gn_apply|gn_temporal' -o gpurun_out/ncu_gn_r3t -f python profiles/gn_ncu_probe.py > gpurun_out/r3t_ncu.log 2>&1
tail -3 gpurun_out/r3t_ncu.log; tail -3 gpurun_out/r3t_plain.log; ls -la gpurun_out/*.ncu-rep
