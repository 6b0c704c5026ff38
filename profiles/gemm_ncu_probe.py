#!/usr/bin/env python
"""Three representative gemm_tc launches for an ncu capture: python profiles/gemm_ncu_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'
N_IMG = 160
for name, H, W, C1, N, taps, use_res, bf16_out in [('conv64_128_128', 64, 64, 128, 128, 9, False, False),
                                                     ('qkv16', 16, 16, 384, 1152, 1, False, True),
                                                     ('proj16', 16, 16, 384, 384, 1, True, False)]:
    M = N_IMG * H * W
    a1 = torch.randn(M, C1, device=dev).bfloat16()
    w = (torch.randn(N, taps * C1, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    resid = torch.randn(M, N, device=dev) if use_res else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if bf16_out else torch.float32)
    st = torch.zeros(N_IMG, 2, N, device=dev, dtype=torch.int64)
    geo = dict(n_img=N_IMG, H=H, W=W)
    for _ in range(3):
        if bf16_out:
            ops.gemm(a1, w, N, taps=taps, bias=bias, residual=resid, out_bf16=out, **geo)
        else:
            ops.gemm(a1, w, N, taps=taps, bias=bias, residual=resid, out_f32=out, stats_out=st, **geo)
    torch.cuda.synchronize()
print('done')
