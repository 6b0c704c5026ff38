#!/usr/bin/env python
"""One launch pair of the output-head conv at the C2 shape for `ncu --set full -k regex:conv3x3_head`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

n, H, W, C, N = 160, 64, 64, 128, 3
M = n * H * W
x = torch.randn(M, C, device='cuda').half()
xb = x.bfloat16()
w = (torch.randn(N, 9 * C, device='cuda') * 0.03).bfloat16()
b = torch.zeros(N, device='cuda')
coef = torch.rand(n, C, 2, device='cuda') + 0.5
out = torch.empty(n, N, H, W, device='cuda')
for _ in range(2):
    ops.gemm(x, w, N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=out, out_nchw=True, a1_coef=coef, a1_act=True)
    ops.gemm(xb, w, N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=out, out_nchw=True)
torch.cuda.synchronize()
