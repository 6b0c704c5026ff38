#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python profiles/linear_ncu_probe.py > gpurun_out/r3p_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_astat --launch-skip 2 -c 2 -o gpurun_out/ncu_linears_r3p -f python profiles/linear_ncu_probe.py > gpurun_out/r3p_ncu.log 2>&1
tail -3 gpurun_out/r3p_ncu.log; ls -la gpurun_out/*.ncu-rep
