#!/usr/bin/env python
"""The two attention linears at the 16x16 level (qkv: bf16 out, TMA-store epilogue; proj_out: residual + fp16 out +
statistics), a few launches each, for `ncu --set full -k regex:gemm_tc_astat` (source-level stall samples)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'
M, C = 40960, 384
a = torch.randn(M, C, device=dev).half()
w16 = (torch.randn(3 * C, C, device=dev) * 0.05).half()
bias3 = torch.zeros(3 * C, device=dev)
qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
att = torch.randn(M, C, device=dev).bfloat16()
wp = (torch.randn(C, C, device=dev) * 0.05).bfloat16()
bias = torch.zeros(C, device=dev)
res = torch.randn(M, C, device=dev).half()
out = torch.empty(M, C, device=dev, dtype=torch.float16)
st = torch.zeros(160, 2, C, device=dev, dtype=torch.int64)
for _ in range(3):
    ops.gemm(a, w16, 3 * C, n_img=M, H=1, W=1, taps=1, bias=bias3, out_bf16=qkv)
    ops.gemm(att, wp, C, n_img=160, H=16, W=16, taps=1, bias=bias, residual=res, out_f32=out, stats_out=st)
torch.cuda.synchronize()
print('ok')

# (--convs) two convolutions of the step as well: 8x8 512->512 (interleaved halo pair kernel, 80 pair tiles on 74 pairs)
# and 64x64 128->128 (transposed-role kernel, the epilogue-heavy K = 1152 shape)
if '--convs' in sys.argv:
    for (n, H, Cc, N) in ((160, 8, 512, 512), (160, 64, 128, 128)):
        x = torch.randn(n * H * H, Cc, device=dev).bfloat16()
        w = (torch.randn(N, 9 * Cc, device=dev) * 0.02).bfloat16()
        b = torch.zeros(N, device=dev)
        o16 = torch.empty(n * H * H, N, device=dev, dtype=torch.float16)
        stc = torch.zeros(n, 2, N, device=dev, dtype=torch.int64)
        for _ in range(2):
            ops.gemm(x, w, N, n_img=n, H=H, W=H, taps=9, bias=b, out_f32=o16, stats_out=stc)
    torch.cuda.synchronize()
    print('convs ok')
