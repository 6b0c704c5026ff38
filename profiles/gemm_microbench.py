#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 implicit-GEMM kernel on the C2 layer shapes (CUDA events,
L2 flushed between launches).  Usage: python profiles/gemm_microbench.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

N_IMG = 160
SHAPES = [  # name, H, W, C1, N, taps, C2, residual
    ('conv64_128_128', 64, 64, 128, 128, 9, 0, True),
    ('lin64_128_128(epilogue probe)', 64, 64, 128, 128, 1, 0, True),
    ('conv64_256_128+skip', 64, 64, 128, 128, 9, 256, False),
    ('conv64_384_128', 64, 64, 384, 128, 9, 0, False),
    ('conv32_256_256', 32, 32, 256, 256, 9, 0, True),
    ('conv32_640_256', 32, 32, 640, 256, 9, 0, False),
    ('conv16_384_384', 16, 16, 384, 384, 9, 0, True),
    ('conv16_896_384', 16, 16, 896, 384, 9, 0, False),
    ('conv8_512_512', 8, 8, 512, 512, 9, 0, True),
    ('conv8_1024_512', 8, 8, 1024, 512, 9, 0, False),
    ('qkv16_384_1152', 16, 16, 384, 1152, 1, 0, False),
    ('proj16_384_384', 16, 16, 384, 384, 1, 0, True),
    ('qkv8_512_1536', 8, 8, 512, 1536, 1, 0, False),
]


def main():
    dev = 'cuda'
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    res = []
    for name, H, W, C1, N, taps, C2, use_res in SHAPES:
        M = N_IMG * H * W
        a1 = torch.randn(M, C1, device=dev).bfloat16()
        a2 = torch.randn(M, C2, device=dev).bfloat16() if C2 else None
        w = (torch.randn(N, taps * C1 + C2, device=dev) * 0.02).bfloat16()
        bias = torch.randn(N, device=dev)
        resid = torch.randn(M, N, device=dev) if use_res else None
        out = torch.empty(M, N, device=dev)
        outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        geo = dict(n_img=N_IMG, H=H, W=W) if taps == 9 else dict(n_img=M, H=1, W=1)
        for variant, kw in (('f32', dict(out_f32=out)), ('bf16', dict(out_bf16=outb))):
            times = []
            for it in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gemm(a1, w, N, taps=taps, a2=a2, bias=bias, residual=resid, **geo, **kw)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    times.append(e0.elapsed_time(e1))
            ms = sorted(times)[len(times) // 2]
            fl = 2.0 * M * N * (taps * C1 + C2)
            res.append(dict(shape=name, out=variant, ms=ms, tflops=fl / ms / 1e9))
            print(f'{name:32s} {variant:5s} {ms * 1e3:9.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s', flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
