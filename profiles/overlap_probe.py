#!/usr/bin/env python
"""How well does an HBM-bound kernel (gn_apply) run BESIDE a persistent tcgen05 conv GEMM of the other micro-batch?
Two streams run the same [gn_apply, conv] sequence on half a batch each, offset by one kernel -- the steady state of
`micro_batches = 2` -- against the same work on one stream.  python profiles/overlap_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402

dev = 'cuda'
REPS = 12


def make(n_img, H, W, C, N):
    M = n_img * H * W
    h = torch.randn(M, C, device=dev).bfloat16()
    a = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(N, 9 * C, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st_in = torch.zeros(n_img, 2, C, device=dev, dtype=torch.int64)
    st_in[:, 0] = 0
    st_in[:, 1] = int(H * W * 16777216)
    st = torch.zeros(n_img, 2, N, device=dev, dtype=torch.int64)
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)

    def E():
        ops.gn_apply(h, None, n_img, H, W, a, stats1=st_in, gamma=gam, beta=bet, silu=True)

    def G():
        ops.gemm(a, w, N, n_img=n_img, H=H, W=W, taps=9, bias=bias, out_bf16=out, stats_out=st)
    return E, G


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


def main():
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, H, W, C, N in [('64x64 128->128', 64, 64, 128, 128), ('32x32 256->256', 32, 32, 256, 256),
                             ('16x16 384->384', 16, 16, 384, 384), ('8x8 512->512', 8, 8, 512, 512)]:
        EA, GA = make(80, H, W, C, N)
        EB, GB = make(80, H, W, C, N)
        for f in (EA, GA, EB, GB):
            f()
        tE = timed(lambda: [EA() for _ in range(REPS)]) / REPS
        tG = timed(lambda: [GA() for _ in range(REPS)]) / REPS

        def serial():
            for _ in range(REPS):
                EA(); GA(); EB(); GB()

        def split():
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)
            s1.wait_event(ev); s2.wait_event(ev)
            with torch.cuda.stream(s1):
                for _ in range(REPS):
                    EA(); GA()
            with torch.cuda.stream(s2):
                for _ in range(REPS):
                    EB(); GB()
            d1, d2 = torch.cuda.Event(), torch.cuda.Event()
            d1.record(s1); d2.record(s2)
            cur.wait_event(d1); cur.wait_event(d2)
        serial(); split()
        tS = timed(serial) / REPS
        tP = timed(split) / REPS
        print(f'{name:16s} half-batch: gn_apply {tE:6.1f} us  conv {tG:6.1f} us | both halves serial {tS:6.1f} us, '
              f'two streams {tP:6.1f} us (pure GEMM time {2 * tG:6.1f} us; hidden {100 * (tS - tP) / (2 * tE):4.0f} % of gn_apply)',
              flush=True)


if __name__ == '__main__':
    print('VDM_GN_THREADS =', os.environ.get('VDM_GN_THREADS', '256 (default)'))
    main()


def diag():
    """Per-stream completion times: GEMMs only on s1, gn_apply only on s2, started together."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, H, W, C, N in [('64x64 128->128', 64, 64, 128, 128), ('32x32 256->256', 32, 32, 256, 256)]:
        EA, GA = make(80, H, W, C, N)
        EB, GB = make(80, H, W, C, N)
        for f in (EA, GA, EB, GB):
            f()
        tE = timed(lambda: [EA() for _ in range(REPS)]) / REPS
        tG = timed(lambda: [GA() for _ in range(REPS)]) / REPS
        for n_e in (REPS, 2 * REPS, 4 * REPS):
            torch.cuda.synchronize()
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event(enable_timing=True)
            d1, d2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev.record(cur)
            s1.wait_event(ev); s2.wait_event(ev)
            with torch.cuda.stream(s1):
                for _ in range(REPS):
                    GA()
                d1.record(s1)
            with torch.cuda.stream(s2):
                for _ in range(n_e):
                    EB()
                d2.record(s2)
            torch.cuda.synchronize()
            print(f'{name}: alone G {tG:.1f} E {tE:.1f} us | together: {REPS} GEMMs done after {ev.elapsed_time(d1) * 1e3:.0f} us '
                  f'(alone {REPS * tG:.0f}), {n_e} gn_apply done after {ev.elapsed_time(d2) * 1e3:.0f} us (alone {n_e * tE:.0f})',
                  flush=True)


if __name__ == '__main__' and os.environ.get('OVERLAP_DIAG'):
    diag()
