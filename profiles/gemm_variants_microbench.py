#!/usr/bin/env python
"""Tile-configuration sweep of the tcgen05 GEMM on the C2 layer shapes (CUDA events, L2 flushed between launches).
VDM_GEMM_MSUB / VDM_GEMM_CTA2 override the dispatch heuristic per call.  Usage: python profiles/gemm_variants_microbench.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

N_IMG = 160
SHAPES = [  # name, H, W, C1, N, taps, C2, residual, bf16 out, stats
    ('qkv16_384_1152', 16, 16, 384, 1152, 1, 0, False, True, False),
    ('proj16_384_384', 16, 16, 384, 384, 1, 0, True, False, True),
    ('qkv8_512_1536', 8, 8, 512, 1536, 1, 0, False, True, False),
    ('proj8_512_512', 8, 8, 512, 512, 1, 0, True, False, True),
    ('conv64_128_128', 64, 64, 128, 128, 9, 0, True, False, True),
    ('conv64_128_128_bf16', 64, 64, 128, 128, 9, 0, False, True, True),
    ('conv64_256_128', 64, 64, 256, 128, 9, 0, False, True, True),
    ('conv64_128_128+skip256', 64, 64, 128, 128, 9, 256, False, False, True),
    ('conv16_384_384', 16, 16, 384, 384, 9, 0, True, False, True),
    ('conv8_512_512', 8, 8, 512, 512, 9, 0, True, False, True),
    ('conv8_1024_512', 8, 8, 1024, 512, 9, 0, False, True, True),
]
VARIANTS = [('auto', {}), ('1cta_msub1', dict(VDM_GEMM_CTA2='0', VDM_GEMM_MSUB='1')),
            ('1cta_msub2', dict(VDM_GEMM_CTA2='0', VDM_GEMM_MSUB='2')), ('2cta', dict(VDM_GEMM_CTA2='2')),
            ('2cta_msub1', dict(VDM_GEMM_CTA2='2', VDM_GEMM_MSUB='1')), ('2cta_msub2', dict(VDM_GEMM_CTA2='2', VDM_GEMM_MSUB='2'))]


def main():
    dev = 'cuda'
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    res = []
    for name, H, W, C1, N, taps, C2, use_res, bf16_out, stats in SHAPES:
        M = N_IMG * H * W
        a1 = torch.randn(M, C1, device=dev).bfloat16()
        a2 = torch.randn(M, C2, device=dev).bfloat16() if C2 else None
        w = (torch.randn(N, taps * C1 + C2, device=dev) * 0.02).bfloat16()
        bias = torch.randn(N, device=dev)
        resid = torch.randn(M, N, device=dev) if use_res else None
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if bf16_out else torch.float32)
        st = torch.zeros(N_IMG, 2, N, device=dev, dtype=torch.int64) if stats else None
        kw = dict(out_bf16=out) if bf16_out else dict(out_f32=out)
        for vname, env in VARIANTS:
            for k in ('VDM_GEMM_CTA2', 'VDM_GEMM_MSUB'):
                os.environ.pop(k, None)
            os.environ.update(env)
            times = []
            for it in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gemm(a1, w, N, taps=taps, a2=a2, bias=bias, residual=resid, n_img=N_IMG, H=H, W=W, stats_out=st, **kw)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    times.append(e0.elapsed_time(e1))
            ms = sorted(times)[len(times) // 2]
            fl = 2.0 * M * N * (taps * C1 + C2)
            res.append(dict(shape=name, variant=vname, ms=ms, tflops=fl / ms / 1e9))
            print(f'{name:24s} {vname:12s} {ms * 1e3:9.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s', flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
