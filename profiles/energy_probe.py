#!/usr/bin/env python
"""Board power of the step's kernel classes: each kernel is replayed from a CUDA graph for ~1.5 s while nvidia-smi
samples power.draw, which gives energy per launch (J) next to the time per launch -- the step as a whole runs at the
1000 W cap, so Joules, not microseconds, are what it is short of.  python profiles/energy_probe.py"""
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_diffusion_b200 import ops  # noqa: E402
from video_diffusion_b200.unet import fold_upsample_weights  # noqa: E402

dev = 'cuda'
n = 160


class Power:
    def __enter__(self):
        self.rows = []
        self.p = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=power.draw,clocks.sm', '--format=csv,noheader,nounits',
                                   '-lms', '50'], stdout=subprocess.PIPE, text=True)
        self.t = threading.Thread(target=lambda: [self.rows.append(l.split(',')) for l in self.p.stdout], daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.p.terminate()
        self.t.join(timeout=2)

    def stats(self):
        rows = self.rows[len(self.rows) // 3:]          # steady state: drop the ramp
        pw = sorted(float(r[0]) for r in rows)
        ck = sorted(int(r[1]) for r in rows)
        return pw[len(pw) // 2], ck[len(ck) // 2]


def measure(name, fn, flops=0.0, nbytes=0.0, reps=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    burst = e0.elapsed_time(e1) * 1e3 / reps
    n_rep = max(4, int(1.6e6 / (burst * reps)))
    with Power() as pw:
        e0.record()
        for _ in range(n_rep):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / (n_rep * reps)
    w, clk = pw.stats()
    extra = (f' {flops / t / 1e6:6.0f} TF/s {flops / 1e12 / (w * t * 1e-6):5.2f} TFLOP/J' if flops else '') + \
            (f' {nbytes / t / 1e3:6.0f} GB/s' if nbytes else '')
    print(f'{name:44s} burst {burst:7.1f} us  sustained {t:7.1f} us  {w:6.0f} W  {clk:5d} MHz  {w * t * 1e-3:7.2f} mJ{extra}',
          flush=True)


def conv(H, W, C, N, res=False, identity=False):
    M = n * H * W
    a = torch.randn(M, C, device=dev).bfloat16()
    w = (torch.randn(N, 9 * C, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev)
    st = torch.zeros(n, 2, N, device=dev, dtype=torch.int64)
    out = torch.empty(M, N, device=dev, dtype=torch.float16 if res else torch.bfloat16)
    r = torch.randn(M, N, device=dev).half() if res else None
    if res:
        return lambda: ops.gemm(a, w, N, n_img=n, H=H, W=W, taps=9, bias=bias, residual=r, out_f32=out, stats_out=st)
    return lambda: ops.gemm(a, w, N, n_img=n, H=H, W=W, taps=9, bias=bias, out_bf16=out, stats_out=st)


def linear(M, N, K, res, HW=256):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.zeros(N, device=dev)
    if not res:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        return lambda: ops.gemm(a, w, N, n_img=M, H=1, W=1, taps=1, bias=bias, out_bf16=out)
    r = torch.randn(M, N, device=dev).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    st = torch.zeros(M // HW, 2, N, device=dev, dtype=torch.int64)
    H = int(HW ** 0.5)
    return lambda: ops.gemm(a, w, N, n_img=M // HW, H=H, W=H, taps=1, bias=bias, residual=r, out_f32=out, stats_out=st)


def main():
    print('idle:', subprocess.run(['nvidia-smi', '--query-gpu=power.draw', '--format=csv,noheader'], capture_output=True,
                                   text=True).stdout.strip())
    for name, H, W, C, N, res in [('conv 64x64 128->128 (K=1152)', 64, 64, 128, 128, False),
                                  ('conv 64x64 256->128 (K=2304)', 64, 64, 256, 128, False),
                                  ('conv 32x32 256->256', 32, 32, 256, 256, False),
                                  ('conv 32x32 256->256 + residual', 32, 32, 256, 256, True),
                                  ('conv 16x16 384->384', 16, 16, 384, 384, False),
                                  ('conv 8x8 512->512', 8, 8, 512, 512, False)]:
        measure(name, conv(H, W, C, N, res), flops=2.0 * n * H * W * N * 9 * C)
    x = torch.randn(n * 32 * 32, 256, device=dev).bfloat16()
    wf = fold_upsample_weights(torch.randn(256, 256, 3, 3) * 0.02).to(dev).bfloat16()
    outu = torch.empty(n * 64 * 64, 256, device=dev, dtype=torch.float16)
    stu = torch.zeros(n, 2, 256, device=dev, dtype=torch.int64)
    bu = torch.zeros(256, device=dev)
    measure('folded upsample 32->64, 256 ch', lambda: ops.gemm(x, wf, 256, n_img=n, H=64, W=64, taps=4, a1_mode=3, bias=bu,
                                                               out_f32=outu, stats_out=stu, C1=256),
            flops=2.0 * n * 4096 * 256 * 4 * 256)
    measure('qkv 16x16', linear(40960, 1152, 384, False), flops=2.0 * 40960 * 1152 * 384)
    measure('proj_out 16x16', linear(40960, 384, 384, True), flops=2.0 * 40960 * 384 * 384)
    measure('qkv 8x8', linear(10240, 1536, 512, False), flops=2.0 * 10240 * 1536 * 512)
    measure('proj_out 8x8', linear(10240, 512, 512, True, 64), flops=2.0 * 10240 * 512 * 512)
    for HWs, C in ((64, 128), (32, 256)):
        M = n * HWs * HWs
        xx = torch.randn(M, C, device=dev).bfloat16()
        out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(n, 2, C, device=dev, dtype=torch.int64)
        st[:, 1] = HWs * HWs * 2 ** 24
        g_, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        measure(f'gn_apply {HWs}x{HWs} C={C}', lambda: ops.gn_apply(xx, None, n, HWs, HWs, out, stats1=st, gamma=g_, beta=b_,
                                                                      silu=True), nbytes=M * C * 4.0)
    qkv = torch.randn(40960, 1152, device=dev).bfloat16()
    att = torch.empty(40960, 384, device=dev, dtype=torch.bfloat16)
    measure('attn_spatial 16x16', lambda: ops.attn_spatial(qkv, n, 256, 4, 96, att), flops=4.0 * n * 4 * 256 * 256 * 96)


if __name__ == '__main__':
    main()
