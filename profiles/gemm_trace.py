#!/usr/bin/env python
"""Where does a tcgen05 GEMM launch spend its time?  Per-CTA cycle counters written by the kernel itself
(vdm_gemm_set_trace): TMA producer waiting for a free smem slot, MMA thread waiting for operands / for a drained
accumulator, epilogue warp waiting for a finished accumulator.  Needs the diagnostics build of the library:
    make -C video_diffusion_b200/csrc clean && make -C video_diffusion_b200/csrc TRACE=1
(rebuild without TRACE afterwards).  TRACE_EXPERIMENTS=1 adds the no-load / no-MMA timing experiments.
Usage: python profiles/gemm_trace.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import _lib, ops  # noqa: E402

N_IMG = 160
SHAPES = [  # name, H, W, C1, N, taps, C2, residual, bf16 out, stats
    ('conv64_128_128_bf16', 64, 64, 128, 128, 9, 0, False, True, True),
    ('conv64_128_128_f32res', 64, 64, 128, 128, 9, 0, True, False, True),
    ('conv64_256_128', 64, 64, 256, 128, 9, 0, False, True, True),
    ('lin64_1152_128', 64, 64, 1152, 128, 1, 0, False, True, True),
    ('conv32_256_256', 32, 32, 256, 256, 9, 0, False, True, True),
    ('qkv16_384_1152', 16, 16, 384, 1152, 1, 0, False, True, False),
    ('proj16_384_384', 16, 16, 384, 384, 1, 0, True, False, True),
    ('conv8_512_512', 8, 8, 512, 512, 9, 0, True, False, True),
]
SHAPES += [('qkv8_512_1536', 8, 8, 512, 1536, 1, 0, False, True, False), ('proj8_512_512', 8, 8, 512, 512, 1, 0, True, False, True)]
if os.environ.get('TRACE_SHAPES'):          # e.g. TRACE_SHAPES=qkv16,proj16,qkv8
    SHAPES = [s_ for s_ in SHAPES if any(s_[0].startswith(k) for k in os.environ['TRACE_SHAPES'].split(','))]
VARIANTS = [('auto', {}), ('1cta_msub1', dict(VDM_GEMM_CTA2='0', VDM_GEMM_MSUB='1')),
            ('1cta_msub2', dict(VDM_GEMM_CTA2='0', VDM_GEMM_MSUB='2')), ('2cta', dict(VDM_GEMM_CTA2='2'))]
if os.environ.get('TRACE_EXPERIMENTS'):     # timing experiments: VDM_GEMM_DEBUG 1 = no TMA loads, 2 = no MMAs
    VARIANTS = [(f'{n}{sfx}', dict(e, **d)) for n, e in VARIANTS for sfx, d in
                (('', {}), ('-noload', dict(VDM_GEMM_DEBUG='1')), ('-nomma', dict(VDM_GEMM_DEBUG='2')))]


def main():
    dev = 'cuda'
    lib = _lib.load()
    lib.vdm_gemm_set_trace.argtypes = [ctypes.c_void_p]
    lib.vdm_gemm_set_trace.restype = None
    trace = torch.zeros(296 * 8, device=dev, dtype=torch.int64)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    print(f'{"shape":24s} {"variant":17s} {"us":>7s} {"TF/s":>7s} | per-CTA mean kcycles: {"mma_total":>9s} {"mma_wait_ops":>12s} '
          f'{"mma_wait_acc":>12s} {"prod_wait":>9s} {"epi_total":>9s} {"epi_wait":>8s}')
    for name, H, W, C1, N, taps, C2, use_res, bf16_out, stats in SHAPES:
        M = N_IMG * H * W
        a1 = torch.randn(M, C1, device=dev).bfloat16()
        w = (torch.randn(N, taps * C1 + C2, device=dev) * 0.02).bfloat16()
        bias = torch.randn(N, device=dev)
        resid = torch.randn(M, N, device=dev) if use_res else None
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if bf16_out else torch.float32)
        st = torch.zeros(N_IMG, 2, N, device=dev, dtype=torch.int64) if stats else None
        kw = dict(out_bf16=out) if bf16_out else dict(out_f32=out)
        geo = dict(n_img=N_IMG, H=H, W=W) if taps == 9 else dict(n_img=M, H=1, W=1)
        if taps == 1 and stats:
            geo = dict(n_img=N_IMG, H=H, W=W)
        for vname, env in VARIANTS:
            for k in ('VDM_GEMM_CTA2', 'VDM_GEMM_MSUB', 'VDM_GEMM_DEBUG'):
                os.environ.pop(k, None)
            os.environ['VDM_GEMM_ASTAT'] = '0'      # the A-stationary kernel carries no counters
            os.environ.update(env)
            times = []
            for it in range(5):
                flush.zero_()
                trace.zero_()
                lib.vdm_gemm_set_trace(trace.data_ptr() if it == 4 else None)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gemm(a1, w, N, taps=taps, bias=bias, residual=resid, stats_out=st, **geo, **kw)
                e1.record()
                torch.cuda.synchronize()
                if 1 <= it < 4:
                    times.append(e0.elapsed_time(e1))
            lib.vdm_gemm_set_trace(None)
            ms = sorted(times)[len(times) // 2]
            t = trace.view(-1, 8).double()
            act = t[t[:, 5] > 0]                      # CTAs that ran an epilogue
            lead = t[t[:, 3] > 0]                     # CTAs that issued MMAs (leaders only in 2-CTA mode)
            m = lambda x: float(x.mean()) / 1e3 if x.numel() else float('nan')
            fl = 2.0 * M * N * (taps * C1 + C2)
            print(f'{name:24s} {vname:17s} {ms * 1e3:7.1f} {fl / ms / 1e9:7.1f} | {"":21s} {m(lead[:, 3]):9.1f} {m(lead[:, 2]):12.1f} '
                  f'{m(lead[:, 1]):12.1f} {m(act[:, 0]):9.1f} {m(act[:, 5]):9.1f} {m(act[:, 4]):8.1f}', flush=True)


if __name__ == '__main__':
    main()
