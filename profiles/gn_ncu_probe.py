#!/usr/bin/env python
"""GroupNorm-apply launches of the step for `ncu --set full -k regex:gn_apply|gn_temporal` (stall samples per SASS
instruction): 64x64 fp16 stream -> bf16 operand (HBM-bound), 16x16 with the stream copy as only output, 8x8 bf16 -> bf16
(latency-bound), and the temporal GroupNorm at 16x16."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'
n = 160


def stats(n, HW, C):
    st = (torch.rand(n, 2, C, device=dev) * 2 ** 24 * HW).long()
    st[:, 1] += st[:, 0].abs() * 4
    return st


for (H, C, in_dt, mode) in ((64, 128, torch.float16, 'a'), (16, 384, torch.float16, 'copy'), (8, 512, torch.bfloat16, 'a')):
    x = torch.randn(n * H * H, C, device=dev).to(in_dt)
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    st = stats(n, H * H, C)
    if mode == 'a':
        out = torch.empty(n * H * H, C, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.gn_apply(x, None, n, H, H, out, stats1=st, gamma=g, beta=b, silu=True)
    else:
        cp = torch.empty_like(x)
        fn = lambda: ops.gn_apply(x, None, n, H, H, None, stats1=st, gamma=g, beta=b, copy=cp)
    for _ in range(3):
        fn()
xt = torch.randn(8, 20, 256, 384, device=dev).half()
r = torch.empty_like(xt)
for _ in range(3):
    ops.gn_temporal(xt, 8, 20, 256, 384, torch.randn(384, device=dev), torch.randn(384, device=dev), r, None)
torch.cuda.synchronize()
print('ok')
