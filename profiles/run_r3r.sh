#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python profiles/linear_ncu_probe.py --convs > gpurun_out/r3r_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 4 -c 6 -o gpurun_out/ncu_gemms_r3r -f python profiles/linear_ncu_probe.py --convs > gpurun_out/r3r_ncu.log 2>&1
tail -3 gpurun_out/r3r_ncu.log; ls -la gpurun_out/*.ncu-rep
