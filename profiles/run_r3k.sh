#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/proj_epilogue_pieces.py > gpurun_out/proj_epilogue_pieces_r3k.log 2>&1; cat gpurun_out/proj_epilogue_pieces_r3k.log | tail -40
