#!/usr/bin/env python
"""Representative tcgen05 GEMM launches of the C2 step for an ncu capture (halo conv kernel at the 64x64 / 32x32 /
16x16 levels (transposed-role kernel for the 128-channel 64x64 layers), its interleaved-tile form at 8x8, the folded-upsample halo kernel, the qkv and proj linears): python profiles/gemm_ncu_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'
N_IMG = 160
for name, H, W, C1, N, taps, use_res, bf16_out in [('conv64_256_128', 64, 64, 256, 128, 9, False, True),
                                                     ('conv64_128_128_res', 64, 64, 128, 128, 9, True, False),
                                                     ('conv128_256_128 (C4 top level, wide-slot kernel)', 128, 128, 256, 128, 9, False, True),
                                                     ('conv32_256_256', 32, 32, 256, 256, 9, False, True),
                                                     ('conv16_384_384', 16, 16, 384, 384, 9, True, False),
                                                     ('conv8_512_512', 8, 8, 512, 512, 9, True, False),
                                                     ('qkv16', 16, 16, 384, 1152, 1, False, True),
                                                     ('proj16', 16, 16, 384, 384, 1, True, False)]:
    M = N_IMG * H * W
    a1 = torch.randn(M, C1, device=dev).bfloat16()
    w = (torch.randn(N, taps * C1, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    # the residual stream of the bf16 model is fp16 (round 2): stream outputs / residuals as the model launches them
    resid = torch.randn(M, N, device=dev).half() if use_res else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if bf16_out else torch.float16)
    st = torch.zeros(N_IMG, 2, N, device=dev, dtype=torch.int64)
    geo = dict(n_img=N_IMG, H=H, W=W)
    for _ in range(2):
        if bf16_out:
            ops.gemm(a1, w, N, taps=taps, bias=bias, residual=resid, out_bf16=out, stats_out=st if taps == 9 else None, **geo)
        else:
            ops.gemm(a1, w, N, taps=taps, bias=bias, residual=resid, out_f32=out, stats_out=st, **geo)
    torch.cuda.synchronize()
# folded nearest-x2 upsample + conv3x3 (low-res 32x32 -> 64x64, 256 -> 256 channels)
from video_diffusion_b200.unet import fold_upsample_weights  # noqa: E402
x = torch.randn(N_IMG * 32 * 32, 256, device=dev).bfloat16()
wf = fold_upsample_weights(torch.randn(256, 256, 3, 3) * 0.02).to(dev).bfloat16()
out = torch.empty(N_IMG * 64 * 64, 256, device=dev, dtype=torch.float16)
st = torch.zeros(N_IMG, 2, 256, device=dev, dtype=torch.int64)
for _ in range(2):
    ops.gemm(x, wf, 256, n_img=N_IMG, H=64, W=64, taps=4, a1_mode=3, bias=torch.zeros(256, device=dev), out_f32=out,
             stats_out=st, C1=256)
torch.cuda.synchronize()
# conv2 of a 64x64 ResBlock as the model runs it: bf16 3x3 part + identity residual read from the fp16 stream (f16 MMAs)
M = N_IMG * 64 * 64
a2n = torch.randn(M, 128, device=dev).bfloat16()
xs = torch.randn(M, 128, device=dev).half()
w3 = (torch.randn(128, 9 * 128, device=dev) * 0.02).bfloat16()
wcat = torch.cat([w3.view(torch.int16), torch.eye(128, device=dev).half().view(torch.int16)], dim=1).contiguous().view(torch.bfloat16)
out = torch.empty(M, 128, device=dev, dtype=torch.float16)
st = torch.zeros(N_IMG, 2, 128, device=dev, dtype=torch.int64)
for _ in range(2):
    ops.gemm(a2n, wcat, 128, n_img=N_IMG, H=64, W=64, taps=9, a2=xs, bias=torch.zeros(128, device=dev), out_f32=out,
             stats_out=st, C1=128)
torch.cuda.synchronize()
print('done')
