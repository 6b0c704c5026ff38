"""Per-kernel parity on the B200: every libvdm entry point against a plain PyTorch fp32
restatement of the same op (TF32 disabled) or against the CPU oracle's sampler maths."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import diffusion_oracle as D  # noqa: E402


@pytest.fixture(scope='module', autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from video_diffusion_b200 import _lib
    _lib.load()          # fail loudly if the extension is missing
    yield


def ops():
    from video_diffusion_b200 import ops as o
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def nhwc(x):   # (n,C,H,W) -> [n*H*W, C]
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def from_nhwc(m, n, h, w):
    return m.view(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


def pack_w(w):  # OIHW -> [O][tap*I + i]
    o, i, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(o, kh * kw * i).contiguous()


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


CONV_CASES = [
    # n, H, W, C1, N, mode, C2
    (2, 64, 64, 64, 128, 0, 0),
    (3, 32, 32, 128, 64, 0, 0),
    (2, 16, 16, 192, 384, 0, 0),
    (5, 8, 8, 128, 128, 0, 0),       # M = 320: ragged last tile
    (10, 4, 4, 64, 64, 0, 0),        # several images per tile, ragged
    (2, 16, 16, 64, 128, 1, 0),      # stride 2 (input 32x32)
    (3, 8, 8, 128, 128, 0, 64),      # fused 1x1 skip operand
    (2, 32, 32, 64, 128, 2, 0),      # upsample folded (materialised for the bf16 kernel)
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['simt_f32', 'tcgen05_bf16'])
@pytest.mark.parametrize('case', CONV_CASES, ids=lambda c: 'n%d_%dx%d_c%d_n%d_m%d_c2_%d' % c)
def test_conv3x3(case, dtype):
    n, H, W, C1, N, mode, C2 = case
    o = ops()
    srcH, srcW = (2 * H, 2 * W) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
    x = rnd(n, C1, srcH, srcW, seed=1)
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5)
    bias = rnd(N, seed=3)
    res = rnd(n * H * W, N, seed=4)
    if dtype == torch.bfloat16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    xin = F.interpolate(x, scale_factor=2, mode='nearest') if mode == 2 else x
    ref = F.conv2d(xin, w, bias, stride=2 if mode == 1 else 1, padding=1)
    wp = pack_w(w)
    a2 = None
    if C2:
        x2 = rnd(n, C2, H, W, seed=5)
        w2 = rnd(N, C2, 1, 1, seed=6, scale=C2 ** -0.5)
        if dtype == torch.bfloat16:
            x2, w2 = x2.bfloat16().float(), w2.bfloat16().float()
        ref = ref + F.conv2d(x2, w2)
        wp = torch.cat([wp, w2.view(N, C2)], dim=1).contiguous()
        a2 = nhwc(x2).to(dtype)
    ref = nhwc(ref) + res
    a1_mode = mode
    if dtype == torch.bfloat16:
        if mode == 1:      # parity planes [n][py][px][H][W][C]
            a1 = x.view(n, C1, H, 2, W, 2).permute(0, 3, 5, 2, 4, 1).contiguous().view(-1, C1)
        elif mode == 2:    # the bf16 kernel consumes the materialised upsample
            a1, a1_mode = nhwc(xin), 0
        else:
            a1 = nhwc(x)
    else:
        a1 = nhwc(x)
    out = torch.empty(n * H * W, N, device='cuda')
    out_b = torch.empty(n * H * W, N, device='cuda', dtype=torch.bfloat16)
    o.gemm(a1.to(dtype), wp.to(dtype), N, n_img=n, H=H, W=W, taps=9, a1_mode=a1_mode, a2=a2, bias=bias,
           residual=res, out_f32=out, out_bf16=out_b, C1=C1)
    torch.cuda.synchronize()
    tol = 2e-5 if dtype == torch.float32 else 2e-5   # bf16 inputs are pre-rounded: only summation order differs
    assert relerr(out, ref) < tol
    assert relerr(out_b, ref) < 6e-3


HALO_CASES = [
    # n, H, W, C1, N, C2, residual
    (3, 64, 64, 128, 128, 0, True),      # 512-row pair tiles, M not a multiple of 512 * k -> ragged pair
    (2, 64, 64, 64, 128, 128, False),    # fused 1x1 skip operand
    (5, 32, 32, 128, 256, 0, False),     # 256-wide pair tiles
    (4, 32, 32, 64, 256, 64, True),
    (7, 16, 16, 192, 384, 0, True),      # 192-wide pair tiles, odd image count
    (4, 16, 16, 64, 192, 0, False),
    (6, 16, 32, 64, 128, 0, False),      # non-square image
    (12, 8, 8, 128, 256, 0, True),       # 8x8: two images per 128-row tile, rows interleaved (y, image, x)
    (7, 8, 8, 64, 512, 128, False),      # 8x8, odd image count, fused skip operand
    (4, 8, 8, 192, 256, 0, False),
    (3, 128, 128, 128, 128, 0, True),    # 128-pixel-wide images: wide-slot transposed kernel, ragged pair (odd count)
    (2, 128, 128, 64, 128, 128, False),  # ... with the fused 1x1 skip operand
    (1, 32, 128, 64, 256, 0, False),     # ... non-square, 256 output channels
]


@pytest.mark.parametrize('case', HALO_CASES, ids=lambda c: 'n%d_%dx%d_c%d_n%d_c2_%d_res%d' % c)
def test_conv3x3_halo_kernel(case, monkeypatch):
    """The halo variant of the 3x3 conv (vertical taps share one activation slot) against conv2d and against the
    plain tcgen05 path (same products, different fp32 summation order), with statistics."""
    n, H, W, C1, N, C2, use_res = case
    o = ops()
    x = rnd(n, C1, H, W, seed=1).bfloat16().float()
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5).bfloat16().float()
    bias = rnd(N, seed=3)
    res = rnd(n * H * W, N, seed=4) if use_res else None
    ref = F.conv2d(x, w, bias, padding=1)
    wp, a2 = pack_w(w), None
    if C2:
        x2 = rnd(n, C2, H, W, seed=5).bfloat16().float()
        w2 = rnd(N, C2, 1, 1, seed=6, scale=C2 ** -0.5).bfloat16().float()
        ref = ref + F.conv2d(x2, w2)
        wp = torch.cat([wp, w2.view(N, C2)], dim=1).contiguous()
        a2 = nhwc(x2).bfloat16()
    ref = nhwc(ref) + (res if use_res else 0)
    outs = []
    for mode in ('2', '0'):
        monkeypatch.setenv('VDM_GEMM_HALO', mode)
        out = torch.full((n * H * W, N), float('nan'), device='cuda')
        st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
        o.gemm(nhwc(x).bfloat16(), wp.bfloat16(), N, n_img=n, H=H, W=W, taps=9, a2=a2, bias=bias, residual=res,
               out_f32=out, stats_out=st, C1=C1)
        assert relerr(out, ref) < 2e-5
        outs.append((out, st))
    assert relerr(outs[0][0], outs[1][0]) < 5e-6
    assert relerr(outs[0][1].double(), outs[1][1].double()) < 1e-5
    # bf16 output variant
    monkeypatch.setenv('VDM_GEMM_HALO', '2')
    out_b = torch.empty(n * H * W, N, device='cuda', dtype=torch.bfloat16)
    o.gemm(nhwc(x).bfloat16(), wp.bfloat16(), N, n_img=n, H=H, W=W, taps=9, a2=a2, bias=bias, residual=res,
           out_bf16=out_b, C1=C1)
    assert relerr(out_b, ref) < 6e-3


@pytest.mark.parametrize('tmode', ['1', '2'], ids=['pair_or_t', 'transposed'])
@pytest.mark.parametrize('case', [c for c in HALO_CASES if c[4] % 128 == 0],
                         ids=lambda c: 'n%d_%dx%d_c%d_n%d_c2_%d_res%d' % c)
def test_conv3x3_fused_groupnorm_silu(case, tmode, monkeypatch):
    """GroupNorm-apply (+ scale/shift) + SiLU fused into the conv's operand path (`a1_coef`: the transform warps of the
    halo kernels rewrite every activation slot in shared memory) against (a) the standalone gn_apply pass feeding the
    same kernel -- BIT-EXACT, the two paths compute the same bf16 operand -- and (b) group_norm + silu + conv2d in
    torch.  Covers the pair, transposed, wide-slot and interleaved 8x8 kernels, the fused 1x1 skip operand (passes
    through untransformed), residuals, ragged pairs and the zero padding (must stay zero after the affine)."""
    n, H, W, C1, N, C2, use_res = case
    o = ops()
    monkeypatch.setenv('VDM_GEMM_HALO', '2')
    monkeypatch.setenv('VDM_GEMM_HALO_T', tmode)
    h1 = (rnd(n, C1, H, W, seed=1) * 1.5 + 0.4).bfloat16()                 # raw conv1 output, bf16 only
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5).bfloat16().float()
    bias = rnd(N, seed=3)
    res = rnd(n * H * W, N, seed=4) if use_res else None
    gamma, beta = 1 + 0.2 * rnd(C1, seed=7), 0.3 * rnd(C1, seed=8)
    ss = rnd(n, 2 * C1, seed=9, scale=0.3)
    wp, a2 = pack_w(w), None
    if C2:
        x2 = rnd(n, C2, H, W, seed=5).bfloat16().float()
        w2 = rnd(N, C2, 1, 1, seed=6, scale=C2 ** -0.5).bfloat16().float()
        wp = torch.cat([wp, w2.view(N, C2)], dim=1).contiguous()
        a2 = nhwc(x2).bfloat16()
    h1r = nhwc(h1.float()).bfloat16()
    st = torch.zeros(n, 2, C1, device='cuda', dtype=torch.float64)
    o.gn_stats(h1r.float(), n, H * W, st)
    kw = dict(n_img=n, H=H, W=W, taps=9, a2=a2, bias=bias, residual=res, C1=C1)
    out_u = torch.full((n * H * W, N), float('nan'), device='cuda')
    st_u = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
    assert o.gemm_fused_norm_supported(h1r, wp.bfloat16(), N, out_f32=out_u, stats_out=st_u, **kw)
    # (a) standalone normalisation pass + the same conv kernel
    a_norm = torch.empty(n * H * W, C1, device='cuda', dtype=torch.bfloat16)
    o.gn_apply(h1r, None, n, H, W, a_norm, stats1=st, gamma=gamma, beta=beta, scale_shift=ss, silu=True)
    o.gemm(a_norm, wp.bfloat16(), N, out_f32=out_u, stats_out=st_u, **kw)
    # fused: raw operand + coefficient table
    coef = torch.empty(n, C1, 2, device='cuda')
    o.gn_coef(st, None, n, H * W, gamma, beta, coef, scale_shift=ss)
    out_f = torch.full((n * H * W, N), float('nan'), device='cuda')
    st_f = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
    o.gemm(h1r, wp.bfloat16(), N, out_f32=out_f, stats_out=st_f, a1_coef=coef, a1_act=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out_f, out_u), float((out_f - out_u).abs().max())
    assert torch.equal(st_f, st_u)
    # (b) torch
    y = F.group_norm(h1.float(), 32, gamma, beta, eps=1e-5) * (1 + ss[:, :C1, None, None]) + ss[:, C1:, None, None]
    ref = F.conv2d(F.silu(y).bfloat16().float(), w, bias, padding=1)
    if C2:
        ref = ref + F.conv2d(x2, w2)
    ref = nhwc(ref) + (res if use_res else 0)
    assert relerr(out_f, ref) < 6e-3          # the operand's bf16 rounding (tanh-form SiLU vs torch's) only
    # bf16 output variant (conv1-style) and affine-only (a1_act = 0)
    out_b = torch.empty(n * H * W, N, device='cuda', dtype=torch.bfloat16)
    st_b = torch.zeros_like(st_f)
    o.gemm(h1r, wp.bfloat16(), N, n_img=n, H=H, W=W, taps=9, a2=a2, bias=bias, out_bf16=out_b, stats_out=st_b,
           a1_coef=coef, a1_act=False, C1=C1)
    refb = F.conv2d(y.bfloat16().float(), w, bias, padding=1)
    if C2:
        refb = refb + F.conv2d(x2, w2)
    assert relerr(out_b, nhwc(refb)) < 8e-3


@pytest.mark.parametrize('M,N,K,n_prob', [(40960, 1152, 384, 1), (10240, 1536, 512, 1), (1000, 384, 384, 1),
                                           (300, 128, 128, 1), (2560, 128, 384, 2), (777, 64, 192, 1)])
def test_linear_tma_store_epilogue(M, N, K, n_prob, monkeypatch):
    """The TMA-store epilogue of the plain linears (bias only; qkv, RPE score GEMMs) against the staged epilogue --
    bit-exact -- and against torch, bf16 and fp32 outputs, ragged M, single-CTA and pair tiles, the two-problem form."""
    o = ops()
    a = rnd(M, K * n_prob, seed=1).bfloat16()
    w = rnd(N * n_prob, K, seed=2, scale=K ** -0.5).bfloat16()
    bias = rnd(N, seed=3) if n_prob == 1 else None
    kw = dict(n_img=M, H=1, W=1, taps=1, bias=bias, C1=K)
    if n_prob > 1:
        kw.update(n_prob=n_prob, prob_a_cols=K, prob_w_rows=N, prob_out_stride=M * N)
    res = {}
    for ts in ('1', '0'):
        monkeypatch.setenv('VDM_GEMM_TS', ts)
        of = torch.full((n_prob, M, N), float('nan'), device='cuda')
        o.gemm(a[:, :K] if n_prob > 1 else a, w, N, out_f32=of, **kw)     # problem i reads column block i of `a`
        ob = None
        if n_prob == 1:
            ob = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
            o.gemm(a, w, N, out_bf16=ob, **kw)
        res[ts] = (of, ob)
    torch.cuda.synchronize()
    assert torch.equal(res['1'][0], res['0'][0])
    if n_prob == 1:
        assert torch.equal(res['1'][1], res['0'][1])
    for i in range(n_prob):
        ref = a[:, i * K:(i + 1) * K].float() @ w[i * N:(i + 1) * N].float().t() + (bias if bias is not None else 0)
        assert relerr(res['1'][0][i], ref) < 2e-5
        if n_prob == 1:
            assert relerr(res['1'][1], ref) < 6e-3


F16_IO_CASES = [
    # n, H, W, C1, N, C2, residual, taps, halo mode
    (3, 64, 64, 128, 128, 0, True, 9, '2'),       # transposed-role kernel
    (2, 64, 64, 64, 128, 128, False, 9, '2'),
    (5, 32, 32, 128, 256, 0, True, 9, '2'),       # pair halo kernel
    (7, 8, 8, 64, 512, 128, False, 9, '2'),       # interleaved 8x8 tiles, fused skip operand
    (12, 8, 8, 128, 256, 0, True, 9, '2'),
    (3, 128, 128, 128, 128, 0, True, 9, '2'),     # wide-slot transposed kernel
    (5, 8, 8, 128, 128, 0, True, 9, '0'),         # plain kernel, ragged tile
    (3, 16, 16, 384, 384, 0, True, 1, '0'),       # attention proj_out: linear with residual + statistics
    (4, 4, 4, 64, 64, 0, True, 9, '0'),           # no fused statistics (H*W % 32 != 0): variants 64 / 65
]


@pytest.mark.parametrize('case', F16_IO_CASES, ids=lambda c: 'n%d_%dx%d_c%d_n%d_c2_%d_res%d_t%d_h%s' % c)
def test_gemm_fp16_stream_io(case, monkeypatch):
    """`io_dtype` = fp16: out_f32 / residual are half tensors (the bf16 model's residual stream).  Accumulation, adds and
    statistics are unchanged fp32, so the result is exactly the fp32-IO result rounded to half, statistics identical."""
    n, H, W, C1, N, C2, use_res, taps, halo = case
    o = ops()
    monkeypatch.setenv('VDM_GEMM_HALO', halo)
    x = nhwc(rnd(n, C1, H, W, seed=1)).bfloat16()
    w = rnd(N, taps * C1 + C2, seed=2, scale=(taps * C1) ** -0.5).bfloat16()
    a2 = nhwc(rnd(n, C2, H, W, seed=5)).bfloat16() if C2 else None
    bias = rnd(N, seed=3)
    res16 = rnd(n * H * W, N, seed=4).half() if use_res else None
    with_stats = (H * W) % 32 == 0
    kw = dict(n_img=n, H=H, W=W, taps=taps, a2=a2, bias=bias, C1=C1)
    o32 = torch.full((n * H * W, N), float('nan'), device='cuda')
    o16 = torch.full((n * H * W, N), float('nan'), device='cuda', dtype=torch.float16)
    st32 = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64) if with_stats else None
    st16 = torch.zeros_like(st32) if with_stats else None
    o.gemm(x, w, N, residual=None if res16 is None else res16.float(), out_f32=o32, stats_out=st32, **kw)
    o.gemm(x, w, N, residual=res16, out_f32=o16, stats_out=st16, **kw)
    torch.cuda.synchronize()
    assert torch.equal(o16, o32.half())
    if with_stats:
        assert torch.equal(st16, st32)
    with pytest.raises(TypeError):           # mixed stream dtypes are refused, not reinterpreted
        o.gemm(x, w, N, residual=rnd(n * H * W, N, seed=4), out_f32=o16, **kw)


A2_F16_CASES = [
    # n, H, W, C1, N, C2a, C2b, halo
    (2, 64, 64, 128, 128, 128, 0, '2'),      # identity residual through the MMA (transposed-role kernel)
    (2, 64, 64, 128, 128, 128, 128, '2'),    # skip projection of a two-tensor concat
    (5, 32, 32, 256, 256, 192, 128, '2'),    # pair halo kernel
    (7, 8, 8, 128, 512, 256, 128, '2'),      # interleaved 8x8 tiles, odd image count
    (3, 128, 128, 128, 128, 128, 64, '2'),   # wide-slot kernel
    (5, 8, 8, 128, 128, 64, 64, '0'),        # plain kernel
    (4, 4, 4, 64, 64, 64, 0, '0'),
]


@pytest.mark.parametrize('case', A2_F16_CASES, ids=lambda c: 'n%d_%dx%d_c%d_n%d_a%d_b%d_h%s' % c)
def test_conv_second_range_from_fp16_stream(case, monkeypatch):
    """The second operand range of a conv (1x1 skip projection / identity residual) read straight from fp16 tensors --
    one or two, concatenated along channels -- against fp16 weight columns (`a2_dtype` = VDM_F16): those K blocks run
    with the f16 MMA format inside the same accumulation as the bf16 3x3 part."""
    n, H, W, C1, N, C2a, C2b, halo = case
    o = ops()
    monkeypatch.setenv('VDM_GEMM_HALO', halo)
    x = rnd(n, C1, H, W, seed=1).bfloat16().float()
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5).bfloat16().float()
    sa, sb = rnd(n, C2a, H, W, seed=3).half(), (rnd(n, C2b, H, W, seed=4).half() if C2b else None)
    identity = C2b == 0 and C2a == N
    ws_ = torch.eye(N, device='cuda') if identity else rnd(N, C2a + C2b, seed=5, scale=(C2a + C2b) ** -0.5)
    ws_ = ws_.half()
    bias = rnd(N, seed=6)
    cat = torch.cat([sa, sb], 1) if C2b else sa
    ref = F.conv2d(x, w, bias, padding=1) + F.conv2d(cat.float(), ws_.float().view(N, -1, 1, 1))
    wp = torch.cat([pack_w(w).bfloat16().view(torch.int16), ws_.view(torch.int16)], dim=1).contiguous().view(torch.bfloat16)
    out = torch.full((n * H * W, N), float('nan'), device='cuda', dtype=torch.float16)
    st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64) if (H * W) % 32 == 0 else None
    o.gemm(nhwc(x).bfloat16(), wp, N, n_img=n, H=H, W=W, taps=9, a2=nhwc(sa.float()).half(),
           a2b=None if sb is None else nhwc(sb.float()).half(), bias=bias, out_f32=out, stats_out=st, C1=C1)
    torch.cuda.synchronize()
    assert relerr(out, nhwc(ref)) < 1e-3            # fp16 output rounding (2^-11) dominates
    if st is not None:
        got = st.double() / 2 ** 24
        assert relerr(got[:, 0], _chan_stats(ref)[:, 0]) < 1e-4


def test_upsample_fold_and_elementwise_fp16_stream(monkeypatch):
    """The folded-upsample conv writing an fp16 stream, and the stream's elementwise consumers / producers
    (gn_stats, gn_apply incl. concat + raw copy + fp16 residual copy, temporal GroupNorm, spatial-encoding add) fed
    with fp16: same results as the fp32 forms on the same (fp16-representable) values."""
    o = ops()
    n, Hl, Wl, C, N = 3, 16, 16, 128, 128
    from video_diffusion_b200.unet import fold_upsample_weights
    x = nhwc(rnd(n, C, Hl, Wl, seed=1)).bfloat16()
    wf = fold_upsample_weights(rnd(N, C, 3, 3, seed=2, scale=(9 * C) ** -0.5)).bfloat16()
    bias = rnd(N, seed=3)
    for halo in ('2', '0'):
        monkeypatch.setenv('VDM_GEMM_HALO', halo)
        o32 = torch.empty(n * 4 * Hl * Wl, N, device='cuda')
        o16 = torch.empty(n * 4 * Hl * Wl, N, device='cuda', dtype=torch.float16)
        s32 = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
        s16 = torch.zeros_like(s32)
        o.gemm(x, wf, N, n_img=n, H=2 * Hl, W=2 * Wl, taps=4, a1_mode=3, bias=bias, out_f32=o32, stats_out=s32, C1=C)
        o.gemm(x, wf, N, n_img=n, H=2 * Hl, W=2 * Wl, taps=4, a1_mode=3, bias=bias, out_f32=o16, stats_out=s16, C1=C)
        assert torch.equal(o16, o32.half()) and torch.equal(s16, s32)
    # elementwise kernels
    n, H, W, C1, C2 = 3, 8, 8, 128, 64
    Cc = C1 + C2
    h1, h2 = rnd(n * H * W, C1, seed=4).half(), rnd(n * H * W, C2, seed=5).half()
    st = [torch.zeros(n, 2, c, device='cuda', dtype=torch.float64) for c in (C1, C2)]
    st_f = [torch.zeros_like(t) for t in st]
    for src, a, b in ((h1, st[0], st_f[0]), (h2, st[1], st_f[1])):
        o.gn_stats(src, n, H * W, a)
        o.gn_stats(src.float(), n, H * W, b)
        assert torch.equal(a, b)
    gamma, beta = rnd(Cc, seed=6), rnd(Cc, seed=7)
    outs = []
    for cast in (lambda t: t, lambda t: t.float()):
        out = torch.empty(n * H * W, Cc, device='cuda', dtype=torch.bfloat16)
        raw = torch.empty_like(out)
        o.gn_apply(cast(h1), cast(h2), n, H, W, out, stats1=st[0], stats2=st[1], gamma=gamma, beta=beta, silu=True,
                   out_raw=raw)
        cp = torch.empty(n * H * W, C1, device='cuda', dtype=cast(h1).dtype)
        out1 = torch.empty(n * H * W, C1, device='cuda', dtype=torch.bfloat16)
        o.gn_apply(cast(h1), None, n, H, W, out1, stats1=st[0], gamma=gamma[:C1].contiguous(), beta=beta[:C1].contiguous(),
                   copy=cp)
        outs.append((out, raw, out1, cp))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert torch.equal(outs[0][3], outs[1][3].half())
    B, T, HW, Ct = 2, 20, 16, 384
    xt = rnd(B, T, HW, Ct, seed=8).half()
    g2, b2 = rnd(Ct, seed=9), rnd(Ct, seed=10)
    r16, a16 = torch.empty_like(xt), torch.empty_like(xt, dtype=torch.bfloat16)
    r32, a32 = torch.empty_like(xt, dtype=torch.float32), torch.empty_like(a16)
    o.gn_temporal(xt, B, T, HW, Ct, g2, b2, r16, a16)
    o.gn_temporal(xt.float(), B, T, HW, Ct, g2, b2, r32, a32)
    assert torch.equal(a16, a32) and torch.equal(r16, r32.half())
    xg = rnd(2, 7, HW, 192, seed=11).half()                      # general (two-pass) temporal kernel
    r16, a16 = torch.empty_like(xg), torch.empty_like(xg, dtype=torch.bfloat16)
    r32, a32 = torch.empty_like(xg, dtype=torch.float32), torch.empty_like(a16)
    o.gn_temporal(xg, 2, 7, HW, 192, g2[:192].contiguous(), b2[:192].contiguous(), r16, a16)
    o.gn_temporal(xg.float(), 2, 7, HW, 192, g2[:192].contiguous(), b2[:192].contiguous(), r32, a32)
    assert torch.equal(a16, a32) and torch.equal(r16, r32.half())
    h = rnd(B * T * HW, Ct, seed=12).half()
    enc, femb = rnd(HW, Ct, seed=13), rnd(B * T, Ct, seed=14)
    e16, e32 = torch.empty_like(h), torch.empty_like(h, dtype=torch.float32)
    o.add_spatial_encoding(h, enc, e16, B * T, HW, Ct, frame_emb=femb)
    o.add_spatial_encoding(h.float(), enc, e32, B * T, HW, Ct, frame_emb=femb)
    assert torch.equal(e16, e32.half())


@pytest.mark.parametrize('n,HW,C,N,kind', [(160, 256, 384, 1152, 'qkv'), (160, 64, 512, 1536, 'qkv'),
                                           (160, 256, 384, 384, 'proj'), (157, 64, 512, 512, 'proj'),
                                           (160, 256, 384, 384, 'proj32'), (40, 256, 128, 256, 'qkv')])
def test_linear_a_stationary_kernel(n, HW, C, N, kind, monkeypatch):
    """The A-stationary pair kernel of the short-K attention linears (a worker keeps the whole K extent of its row tile
    in shared memory and streams only weight tiles) against the plain kernel: same products in the same k order, so
    outputs and GroupNorm statistics are bit-identical; plus torch."""
    o = ops()
    M = n * HW
    H = W = int(HW ** 0.5)
    a = rnd(M, C, seed=1).bfloat16()
    w = rnd(N, C, seed=2, scale=C ** -0.5).bfloat16()
    bias = rnd(N, seed=3)
    outs = {}
    monkeypatch.setenv('VDM_GEMM_LINT', '0')      # (the transposed-role kernel takes these shapes by default)
    for mode in ('1', '0'):
        monkeypatch.setenv('VDM_GEMM_ASTAT', mode)
        if kind == 'qkv':
            out = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
            o.gemm(a, w, N, n_img=M, H=1, W=1, taps=1, bias=bias, out_bf16=out)
            outs[mode] = (out, None)
        else:
            io = torch.float32 if kind == 'proj32' else torch.float16
            res = rnd(M, N, seed=4).to(io)
            out = torch.full((M, N), float('nan'), device='cuda', dtype=io)
            st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
            o.gemm(a, w, N, n_img=n, H=H, W=W, taps=1, bias=bias, residual=res, out_f32=out, stats_out=st)
            outs[mode] = (out, st)
    torch.cuda.synchronize()
    assert torch.equal(outs['1'][0], outs['0'][0])
    ref = a.float() @ w.float().t() + bias
    if kind != 'qkv':
        assert torch.equal(outs['1'][1], outs['0'][1])
        ref = ref + res.float()
    assert relerr(outs['1'][0], ref) < 6e-3


@pytest.mark.parametrize('M,C,N', [(40960, 384, 1152), (10240, 512, 1536), (2560, 128, 256), (1000, 64, 24)])
def test_linear_fp16_operands(M, C, N, monkeypatch):
    """dtype VDM_F16: a plain linear whose activations AND weights are IEEE half (the normalised fp16 stream feeding the
    attention qkv projection), on the A-stationary pair kernel, the plain kernel and the transposed-role kernel --
    bit-identical to each other (same products, same k order), and against torch on the same fp16 values."""
    o = ops()
    a = rnd(M, C, seed=1).half()
    w = rnd(N, C, seed=2, scale=C ** -0.5).half()
    bias = rnd(N, seed=3)
    outs = []
    for astat, lint in (('1', '0'), ('0', '0'), ('0', '2')):
        monkeypatch.setenv('VDM_GEMM_ASTAT', astat)
        monkeypatch.setenv('VDM_GEMM_LINT', lint)
        out = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
        o.gemm(a, w, N, n_img=M, H=1, W=1, taps=1, bias=bias, out_bf16=out)
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = a.float() @ w.float().t() + bias
    assert relerr(outs[0], ref) < 5e-3            # bf16 rounding of the output only
    # the same values as bf16 operands lose three significand bits
    out_b = torch.empty_like(outs[0])
    o.gemm(a.bfloat16(), w.bfloat16(), N, n_img=M, H=1, W=1, taps=1, bias=bias, out_bf16=out_b)
    assert relerr(outs[0], ref) <= relerr(out_b, ref) + 1e-4
    with pytest.raises(RuntimeError, match='plain linears'):          # convolutions keep bf16 operands
        o.gemm(rnd(2 * 64, 64, seed=5).half(), rnd(64, 9 * 64, seed=6).half(), 64, n_img=2, H=8, W=8, taps=9,
               out_bf16=torch.empty(128, 64, device='cuda', dtype=torch.bfloat16))


def test_groupnorm_kernels_with_the_stream_copy_as_only_output():
    """gn_temporal with out_a = None and gn_apply with out = None + copy: only the normalised fp16 stream copy is
    written (it feeds the qkv projection directly); bit-identical to the copy of the two-output call."""
    o = ops()
    B, T, HW, Cc = 2, 20, 64, 384
    x = (rnd(B, T, HW, Cc, seed=1) * 2 + 0.5).half()
    g, b = 1 + 0.1 * rnd(Cc, seed=2), 0.1 * rnd(Cc, seed=3)
    both_r, both_a = torch.empty_like(x), torch.empty(B, T, HW, Cc, device='cuda', dtype=torch.bfloat16)
    o.gn_temporal(x, B, T, HW, Cc, g, b, both_r, both_a)
    only = torch.full_like(x, float('nan'))
    o.gn_temporal(x, B, T, HW, Cc, g, b, only, None)
    assert torch.equal(only, both_r)
    n = B * T
    xs = x.view(n * HW, Cc)
    st = _chan_stats(xs.view(n, 8, 8, Cc).permute(0, 3, 1, 2).float())
    c2, a2 = torch.empty_like(xs), torch.empty(n * HW, Cc, device='cuda', dtype=torch.bfloat16)
    o.gn_apply(xs, None, n, 8, 8, a2, stats1=st, gamma=g, beta=b, copy=c2)
    c1 = torch.full_like(xs, float('nan'))
    o.gn_apply(xs, None, n, 8, 8, None, stats1=st, gamma=g, beta=b, copy=c1)
    assert torch.equal(c1, c2)
    with pytest.raises(RuntimeError, match='NULL'):
        o.gn_apply(xs, None, n, 8, 8, None, stats1=st, gamma=g, beta=b)


@pytest.mark.parametrize('n,HW,C,N,kind', [(160, 256, 384, 1152, 'qkv'), (160, 64, 512, 1536, 'qkv'),
                                           (160, 256, 384, 384, 'proj'), (157, 64, 512, 512, 'proj'),
                                           (80, 256, 384, 384, 'proj32'), (7, 128, 64, 128, 'proj'),
                                           (33, 64, 128, 256, 'proj_nores'), (3, 256, 192, 384, 'qkv')])
def test_linear_on_transposed_role_kernel(n, HW, C, N, kind, monkeypatch):
    """Linears on the transposed-role kernel (weights = M operand, 256 pixels = N operand, every K block through the
    second-range path; lean epilogue with per-lane GroupNorm statistics, incl. two 64-pixel images per warp) against
    the generic kernel and torch.  Ragged row counts, column-block views of A, fp16 / fp32 stream IO."""
    o = ops()
    M = n * HW
    H = W = int(HW ** 0.5) if int(HW ** 0.5) ** 2 == HW else 0
    if H == 0:
        H, W = HW // 8, 8
    wide = rnd(M, C + 64, seed=1).bfloat16()
    a = wide[:, 64:]                                   # a column block of a wider tensor (lda1 > C1)
    w = rnd(N, C, seed=2, scale=C ** -0.5).bfloat16()
    bias = rnd(N, seed=3)
    outs = {}
    monkeypatch.setenv('VDM_GEMM_ASTAT', '0')
    for mode in ('2', '0'):
        monkeypatch.setenv('VDM_GEMM_LINT', mode)
        if kind == 'qkv':
            out = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
            o.gemm(a, w, N, n_img=M, H=1, W=1, taps=1, bias=bias, out_bf16=out)
            outs[mode] = (out, None)
        else:
            io = torch.float32 if kind == 'proj32' else torch.float16
            res = None if kind == 'proj_nores' else rnd(M, N, seed=4).to(io)
            out = torch.full((M, N), float('nan'), device='cuda', dtype=io)
            st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
            o.gemm(a, w, N, n_img=n, H=H, W=W, taps=1, bias=bias, residual=res, out_f32=out, stats_out=st)
            outs[mode] = (out, st)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    if kind != 'qkv':
        if res is not None:
            ref = ref + res.float()
        got = outs['2'][1].double() / 2 ** 24
        want = torch.stack([outs['2'][0].double().view(n, HW, N).sum(1), (outs['2'][0].double() ** 2).view(n, HW, N).sum(1)], 1)
        # statistics of the fp32 values before the output rounding: compare with the generic kernel's and with the
        # rounded output's own sums
        assert relerr(got, outs['0'][1].double() / 2 ** 24) < 1e-5
        assert relerr(got, want) < (1e-5 if kind == 'proj32' else 2e-3)
    assert torch.isfinite(outs['2'][0].float()).all()
    assert relerr(outs['2'][0], outs['0'][0]) < 2e-3
    assert relerr(outs['2'][0], ref) < 6e-3


@pytest.mark.parametrize('n,H,W,C,N', [(160, 64, 64, 128, 128), (80, 16, 16, 384, 384), (7, 32, 32, 64, 128),
                                       (3, 64, 64, 64, 384)])
def test_groupnorm_apply_pipelined_behind_conv(n, H, W, C, N, monkeypatch):
    """conv (transposed-role kernel) publishes per-image completion counters; the GroupNorm-apply launched right behind
    it as a programmatic dependent normalises image by image while the conv still runs.  Bit-identical to the
    serialised pair, counters end at H*W*N, repeated rounds (stale data from the previous round must never be read)."""
    o = ops()
    monkeypatch.setenv('VDM_GEMM_HALO', '2')
    monkeypatch.setenv('VDM_GEMM_HALO_T', '2')
    M = n * H * W
    w = (rnd(N, 9 * C, seed=2) * (9 * C) ** -0.5).bfloat16()
    bias, gam, bet = rnd(N, seed=3), rnd(N, seed=4), rnd(N, seed=5)
    ss = rnd(n, 2 * N, seed=6) * 0.1
    h = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
    done = torch.zeros(n, device='cuda', dtype=torch.int32)
    conv = dict(n_img=n, H=H, W=W, taps=9, bias=bias, out_bf16=h, stats_out=st)
    a0 = rnd(M, C, seed=1).bfloat16()
    assert o.gemm_img_done_supported(a0, w, N, **conv)
    for rnd_i in range(3):
        a = (rnd(M, C, seed=10 + rnd_i)).bfloat16()
        # serialised reference
        st.zero_()
        o.gemm(a, w, N, **conv)
        ref = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
        o.gn_apply(h, None, n, H, W, ref, stats1=st, gamma=gam, beta=bet, scale_shift=ss, silu=True)
        st_ref = st.clone()
        torch.cuda.synchronize()
        # pipelined: poison the outputs, then conv + dependent apply
        h.fill_(float('nan'))
        st.zero_()
        done.zero_()
        out = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
        o.gemm(a, w, N, img_done=done, **conv)
        o.gn_apply(h, None, n, H, W, out, stats1=st, gamma=gam, beta=bet, scale_shift=ss, silu=True, wait_done=done)
        torch.cuda.synchronize()
        assert torch.equal(done, torch.full_like(done, H * W * N))
        assert torch.equal(st, st_ref)
        assert torch.equal(out.view(torch.int16), ref.view(torch.int16))


def test_img_done_unsupported_kernels_raise(monkeypatch):
    o = ops()
    a = rnd(160 * 64, 512, seed=1).bfloat16()
    w = rnd(512, 9 * 512, seed=2).bfloat16()
    out = torch.empty(160 * 64, 512, device='cuda', dtype=torch.bfloat16)
    done = torch.zeros(160, device='cuda', dtype=torch.int32)
    kw = dict(n_img=160, H=8, W=8, taps=9, out_bf16=out)
    assert not o.gemm_img_done_supported(a, w, 512, **kw)
    with pytest.raises(RuntimeError):
        o.gemm(a, w, 512, img_done=done, **kw)


def test_fused_groupnorm_unsupported_shapes_raise():
    """a1_coef on a shape / epilogue the transform-stage kernels do not cover is an error, not a silent slow path."""
    o = ops()
    n, H, W, C1, N = 2, 4, 4, 64, 64
    x = rnd(n * H * W, C1, seed=1).bfloat16()
    w = rnd(N, 9 * C1, seed=2).bfloat16()
    coef = torch.ones(n, C1, 2, device='cuda')
    out = torch.empty(n * H * W, N, device='cuda')
    st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
    assert not o.gemm_fused_norm_supported(x, w, N, n_img=n, H=H, W=W, taps=9, out_f32=out)
    with pytest.raises(RuntimeError):
        o.gemm(x, w, N, n_img=n, H=H, W=W, taps=9, out_f32=out, a1_coef=coef)
    with pytest.raises(RuntimeError):        # linears never take it
        o.gemm(x, w[:, :C1].contiguous(), N, n_img=n * H * W, H=1, W=1, taps=1, out_f32=out, a1_coef=coef)
    del st


@pytest.mark.parametrize('halo', ['0', '2'], ids=['plain', 'halo'])
@pytest.mark.parametrize('case', [(2, 16, 16, 128, 128), (3, 32, 32, 64, 256), (160, 8, 8, 128, 128), (5, 64, 64, 128, 256),
                                  (3, 128, 128, 64, 128), (7, 32, 64, 64, 384), (1, 16, 32, 64, 128), (3, 16, 32, 128, 256)])
def test_upsample_conv_folded_into_parity_convs(case, halo, monkeypatch):
    """nearest-x2 + conv3x3 as four 2x2 convs on the low-res input (a1_mode 3) vs F.interpolate + F.conv2d; `halo`:
    the variant that keeps both vertical parities of a tile as two accumulators sharing every activation box (low-res
    widths 16 / 32 / 64; other shapes fall back to the plain kernel)."""
    from video_diffusion_b200.unet import fold_upsample_weights
    monkeypatch.setenv('VDM_GEMM_HALO', halo)
    o = ops()
    n, H, W, C1, N = case                      # H, W: OUTPUT resolution
    x = rnd(n, C1, H // 2, W // 2, seed=1).bfloat16().float()
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5)
    bias = rnd(N, seed=3)
    wf = fold_upsample_weights(w).cuda().bfloat16()
    # reference with the SAME (folded, bf16-rounded) weights: exact algebra, only summation order differs
    ref = torch.empty(n, N, H, W, device='cuda')
    xp = F.pad(x, (1, 1, 1, 1))
    wf4 = wf.float().view(4, N, 2, 2, C1).permute(0, 1, 4, 2, 3)
    for a in range(2):
        for b in range(2):
            win = xp[:, :, a:a + H // 2 + 1, b:b + W // 2 + 1]
            ref[:, :, a::2, b::2] = F.conv2d(win, wf4[2 * a + b], bias)
    # and the folding itself against the un-folded definition in fp32
    full = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, bias, padding=1)
    wf32 = fold_upsample_weights(w).cuda().view(4, N, 2, 2, C1).permute(0, 1, 4, 2, 3)
    chk = torch.empty_like(full)
    for a in range(2):
        for b in range(2):
            chk[:, :, a::2, b::2] = F.conv2d(xp[:, :, a:a + H // 2 + 1, b:b + W // 2 + 1], wf32[2 * a + b], bias)
    assert relerr(chk, full) < 1e-5
    out = torch.empty(n * H * W, N, device='cuda')
    st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64) if (H // 2) * (W // 2) % 32 == 0 else None
    o.gemm(nhwc(x).bfloat16(), wf, N, n_img=n, H=H, W=W, taps=4, a1_mode=3, bias=bias, out_f32=out, stats_out=st, C1=C1)
    assert relerr(out, nhwc(ref)) < 2e-5
    if st is not None:
        cs = _chan_stats(from_nhwc(out, n, H, W))
        assert relerr(st.double()[:, 0] / 2 ** 24, cs[:, 0]) < 2e-6
        assert relerr(st.double()[:, 1] / 2 ** 24, cs[:, 1]) < 2e-6


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['simt_f32', 'tcgen05_bf16'])
def test_linear_rowbias_nchw_and_silu(dtype):
    o = ops()
    M, K, N = 300, 128, 192
    a = rnd(M, K, seed=1)
    w = rnd(N, K, seed=2, scale=K ** -0.5)
    if dtype == torch.bfloat16:
        a, w = a.bfloat16().float(), w.bfloat16().float()
    bias = rnd(N, seed=3)
    ref = a @ w.t() + bias
    out = torch.empty(M, N, device='cuda')
    sil = torch.empty(M, N, device='cuda') if dtype == torch.float32 else None
    o.gemm(a.to(dtype), w.to(dtype), N, n_img=M, H=1, W=1, taps=1, bias=bias, out_f32=out, out_silu=sil)
    assert relerr(out, ref) < 2e-5
    if sil is not None:
        assert relerr(sil, F.silu(ref)) < 2e-5
    # per-image bias + NCHW store of a 3-channel 3x3 conv (the output head)
    n, H, W, C1 = 3, 16, 16, 64
    x = rnd(n, C1, H, W, seed=5)
    wc = rnd(3, C1, 3, 3, seed=6, scale=(9 * C1) ** -0.5)
    if dtype == torch.bfloat16:
        x, wc = x.bfloat16().float(), wc.bfloat16().float()
    b3 = rnd(3, seed=7)
    ref = F.conv2d(x, wc, b3, padding=1)
    out = torch.empty(n, 3, H, W, device='cuda')
    o.gemm(nhwc(x).to(dtype), pack_w(wc).to(dtype), 3, n_img=n, H=H, W=W, taps=9, bias=b3, out_f32=out, out_nchw=True)
    assert relerr(out, ref) < 2e-5
    rb = rnd(n, 2 * 64, seed=8)
    wq = rnd(64, C1, 3, 3, seed=9, scale=(9 * C1) ** -0.5)
    if dtype == torch.bfloat16:
        wq = wq.bfloat16().float()
    ref = nhwc(F.conv2d(x, wq, None, padding=1) + rb[:, :64, None, None])
    out = torch.empty(n * H * W, 64, device='cuda')
    o.gemm(nhwc(x).to(dtype), pack_w(wq).to(dtype), 64, n_img=n, H=H, W=W, taps=9, rowbias=rb, out_f32=out)
    assert relerr(out, ref) < 2e-5


@pytest.mark.parametrize('n,H,W,C1,N', [(2, 64, 64, 128, 3), (3, 32, 32, 64, 6), (2, 16, 16, 192, 3), (1, 8, 128, 64, 3), (2, 128, 128, 128, 3), (1, 6, 128, 64, 6),
                                       (4, 8, 8, 64, 3), (2, 32, 32, 256, 8), (1, 16, 16, 64, 1)])
def test_output_head_conv_small_n(n, H, W, C1, N, monkeypatch):
    """C -> 3 / 6 channel 3x3 conv with a planar (NCHW) store: the streaming mma.sync kernel (conv_small_n.cu; ragged row tiles, 1 to 8
    output channels) and, for shapes it does not take (8x8 here) or when switched off, the tcgen05 GEMM; both against conv2d on the same bf16-rounded data."""
    o = ops()
    x = rnd(n, C1, H, W, seed=1).bfloat16().float()
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5).bfloat16().float()
    b = rnd(N, seed=3)
    ref = F.conv2d(x, w, b, padding=1)
    for disable in (False, True):
        if disable:
            monkeypatch.setenv('VDM_NO_SMALL_N', '1')
        out = torch.full((n, N, H, W), float('nan'), device='cuda')
        o.gemm(nhwc(x).bfloat16(), pack_w(w).bfloat16(), N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=out,
               out_nchw=True)
        assert relerr(out, ref) < 2e-5


@pytest.mark.parametrize('n,H,W,C1,N', [(1, 37, 64, 128, 3), (1, 20, 16, 64, 1), (2, 24, 48, 64, 2), (1, 5, 80, 192, 8)])
def test_output_head_conv_ragged_shapes(n, H, W, C1, N):
    """Image heights that are no multiple of the row tile, widths that are no power of two (any multiple of 16):
    shapes only the streaming kernel takes (the tcgen05 GEMM wants H*W in multiples of 128)."""
    o = ops()
    x = rnd(n, C1, H, W, seed=1).bfloat16().float()
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5).bfloat16().float()
    b = rnd(N, seed=3)
    out = torch.full((n, N, H, W), float('nan'), device='cuda')
    o.gemm(nhwc(x).bfloat16(), pack_w(w).bfloat16(), N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=out, out_nchw=True)
    assert relerr(out, F.conv2d(x, w, b, padding=1)) < 2e-5


@pytest.mark.parametrize('n,H,W,C1,N', [(2, 64, 64, 128, 3), (3, 32, 32, 64, 6), (2, 16, 16, 192, 3), (1, 8, 128, 64, 3),
                                       (2, 128, 128, 128, 3), (2, 32, 32, 256, 6)])
def test_output_head_conv_with_fused_groupnorm_silu(n, H, W, C1, N):
    """The output head reading the RAW fp16 stream: GroupNorm-apply + SiLU while the tile is staged (a1_coef +
    a1_raw_dtype), against the standalone gn_apply -> head conv pair on the same data, and against torch."""
    o = ops()
    x = (rnd(n, C1, H, W, seed=1) * 1.5 + 0.3).half()
    w = rnd(N, C1, 3, 3, seed=2, scale=(9 * C1) ** -0.5).bfloat16()
    b = rnd(N, seed=3)
    gamma, beta = 1 + 0.1 * rnd(C1, seed=4), 0.1 * rnd(C1, seed=5)
    st = _chan_stats(x.float())
    xs = nhwc(x)
    coef = torch.empty(n, C1, 2, device='cuda')
    o.gn_coef(st, None, n, H * W, gamma, beta, coef)
    out = torch.full((n, N, H, W), float('nan'), device='cuda')
    o.gemm(xs, pack_w(w.float()).bfloat16(), N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=out, out_nchw=True,
           a1_coef=coef, a1_act=True)
    a = torch.empty(n * H * W, C1, device='cuda', dtype=torch.bfloat16)
    o.gn_apply(xs, None, n, H, W, a, stats1=st, gamma=gamma, beta=beta, silu=True)
    two = torch.empty_like(out)
    o.gemm(a, pack_w(w.float()).bfloat16(), N, n_img=n, H=H, W=W, taps=9, bias=b, out_f32=two, out_nchw=True)
    assert relerr(out, two) < 1e-5          # same bf16 operand values (shared silu / affine arithmetic)
    ref = F.conv2d(F.silu(F.group_norm(x.float(), 32, gamma, beta, eps=1e-5)), w.float(), b, padding=1)
    assert relerr(out, ref) < 6e-3          # the activated operand is rounded to bf16


def test_output_head_fused_groupnorm_unsupported_shape_raises():
    o = ops()
    n, H, W, C1, N = 4, 8, 8, 64, 3             # 8-pixel-wide images go to the tcgen05 GEMM, which has no such stage
    x = rnd(n * H * W, C1, seed=1).half()
    coef = torch.ones(n, C1, 2, device='cuda')
    out = torch.empty(n, N, H, W, device='cuda')
    with pytest.raises(RuntimeError, match='fused-normalisation'):
        o.gemm(x, rnd(N, 9 * C1, seed=2).bfloat16(), N, n_img=n, H=H, W=W, taps=9, out_f32=out, out_nchw=True,
               a1_coef=coef, a1_act=True)


def _chan_stats(x, dtype=torch.float64):
    """[n][2][C] per-(image, channel) sum and sum of squares of an NCHW tensor."""
    xd = x.double()
    return torch.stack([xd.sum(dim=(2, 3)), (xd * xd).sum(dim=(2, 3))], dim=1).to(dtype).contiguous()


@pytest.mark.parametrize('C1,C2', [(64, 0), (128, 64), (512, 384), (192, 0)])
@pytest.mark.parametrize('out_dtype', [torch.float32, torch.bfloat16])
def test_groupnorm_apply(C1, C2, out_dtype):
    o = ops()
    n, H, W = 3, 8, 8
    Cc = C1 + C2
    x1 = rnd(n, C1, H, W, seed=1) * 2 + 0.5
    x2 = rnd(n, C2, H, W, seed=2) if C2 else None
    gamma, beta = rnd(Cc, seed=3), rnd(Cc, seed=4)
    ss = rnd(n, 2 * Cc, seed=5, scale=0.3)
    xc = torch.cat([x1, x2], 1) if C2 else x1
    ref = F.group_norm(xc, 32, gamma, beta, eps=1e-5) * (1 + ss[:, :Cc, None, None]) + ss[:, Cc:, None, None]
    ref = F.silu(ref)
    s1, s2 = nhwc(x1), (nhwc(x2) if C2 else None)
    st1 = torch.zeros(n, 2, C1, device='cuda', dtype=torch.float64)
    o.gn_stats(s1, n, H * W, st1)
    assert relerr(st1, _chan_stats(x1)) < 1e-12
    st2 = None
    if C2:
        st2 = torch.zeros(n, 2, C2, device='cuda', dtype=torch.float64)
        o.gn_stats(s2, n, H * W, st2)
    out = torch.empty(n * H * W, Cc, device='cuda', dtype=out_dtype)
    raw = torch.empty(n * H * W, Cc, device='cuda', dtype=out_dtype)
    o.gn_apply(s1, s2, n, H, W, out, stats1=st1, stats2=st2, gamma=gamma, beta=beta, scale_shift=ss, silu=True,
               out_raw=raw)
    tol = 1e-5 if out_dtype == torch.float32 else 5e-3
    assert relerr(out, nhwc(ref)) < tol
    assert relerr(raw, nhwc(xc)) < tol
    # fixed-point per-channel tables, as the bf16 GEMM epilogue produces them
    fx = lambda t: None if t is None else (t * 2 ** 24).round().long()
    o.gn_apply(s1, s2, n, H, W, out, stats1=fx(st1), stats2=fx(st2), gamma=gamma, beta=beta, scale_shift=ss, silu=True)
    assert relerr(out, nhwc(ref)) < max(tol, 2e-5)
    # plain norm + fp32 copy, x2 upsample and parity layouts of the raw cast
    cp = torch.empty(n * H * W, Cc, device='cuda')
    o.gn_apply(s1, s2, n, H, W, out, stats1=st1, stats2=st2, gamma=gamma, beta=beta, copy=cp)
    assert relerr(cp, nhwc(F.group_norm(xc, 32, gamma, beta, eps=1e-5))) < 1e-5
    up = torch.empty(n * 4 * H * W, Cc, device='cuda', dtype=out_dtype)
    o.gn_apply(s1, s2, n, H, W, up, out_mode=1)
    assert relerr(up, nhwc(F.interpolate(xc, scale_factor=2, mode='nearest'))) < 5e-3
    par = torch.empty(n * H * W, Cc, device='cuda', dtype=out_dtype)
    o.gn_apply(s1, s2, n, H, W, par, out_mode=2)
    refp = xc.view(n, Cc, H // 2, 2, W // 2, 2).permute(0, 3, 5, 2, 4, 1).reshape(-1, Cc)
    assert relerr(par, refp) < 5e-3


def test_gemm_epilogue_groupnorm_statistics():
    """The bf16 GEMM's fused per-(image, channel) sums against sums of its own stored output."""
    o = ops()
    for (n, H, W, C1, N, taps) in [(3, 16, 16, 64, 128, 9), (5, 8, 8, 128, 384, 9), (2, 32, 32, 64, 64, 9),
                                   (4, 8, 8, 128, 256, 1), (160, 8, 8, 64, 128, 9)]:
        x = rnd(n, C1, H, W, seed=1).bfloat16()
        w = rnd(N, taps * C1, seed=2, scale=(taps * C1) ** -0.5).bfloat16()
        bias, res = rnd(N, seed=3), rnd(n * H * W, N, seed=4)
        out = torch.empty(n * H * W, N, device='cuda')
        st = torch.zeros(n, 2, N, device='cuda', dtype=torch.int64)
        o.gemm(nhwc(x.float()).bfloat16(), w, N, n_img=n, H=H, W=W, taps=taps, bias=bias, residual=res, out_f32=out,
               stats_out=st)
        got = st.double() / 2 ** 24
        ref = _chan_stats(from_nhwc(out, n, H, W))
        assert relerr(got[:, 0], ref[:, 0]) < 2e-6 and relerr(got[:, 1], ref[:, 1]) < 2e-6
        st2 = torch.zeros_like(st)
        o.gemm(nhwc(x.float()).bfloat16(), w, N, n_img=n, H=H, W=W, taps=taps, bias=bias, residual=res, out_f32=out,
               stats_out=st2)
        assert torch.equal(st, st2)          # integer accumulation: bit-reproducible


@pytest.mark.parametrize('B,T,HW,Cc', [(2, 5, 16, 64), (2, 20, 64, 384), (3, 7, 16, 512), (1, 32, 4, 128),
                                        (2, 20, 16, 640), (2, 33, 4, 128), (2, 6, 16, 192)])
def test_groupnorm_temporal_and_spatial_encoding(B, T, HW, Cc):
    """Both temporal GroupNorm kernels: the register-resident single pass (C % 128 == 0, T <= 32) and the general one."""
    o = ops()
    x = rnd(B, T, HW, Cc, seed=1) + 0.3
    gamma, beta = rnd(Cc, seed=2), rnd(Cc, seed=3)
    xr = x.permute(0, 2, 3, 1).reshape(B * HW, Cc, T)              # (B*D, C, T) as the reference
    ref = F.group_norm(xr, 32, gamma, beta, eps=1e-5).view(B, HW, Cc, T).permute(0, 3, 1, 2)
    out = torch.empty_like(x)
    outa = torch.empty_like(x, dtype=torch.bfloat16)
    o.gn_temporal(x, B, T, HW, Cc, gamma, beta, out, outa)
    assert relerr(out, ref) < 1e-5
    assert relerr(outa, ref) < 5e-3
    h = rnd(B * T * HW, Cc, seed=4)
    enc = rnd(HW, Cc, seed=5)
    ref = (h.view(B * T, HW, Cc) + enc).view(-1, Cc)
    o.add_spatial_encoding(h, enc, h, B * T, HW, Cc)
    assert relerr(h, ref) < 1e-7
    # + per-frame sinusoid (use_frame_encoding, unet.py:914-926); either addend alone as well
    femb = rnd(B * T, Cc, seed=6)
    h2 = rnd(B * T * HW, Cc, seed=7)
    both, only_f = torch.empty_like(h2), torch.empty_like(h2)
    o.add_spatial_encoding(h2, enc, both, B * T, HW, Cc, frame_emb=femb)
    o.add_spatial_encoding(h2, None, only_f, B * T, HW, Cc, frame_emb=femb)
    assert torch.equal(both.view(B * T, HW, Cc), (h2.view(B * T, HW, Cc) + enc) + femb[:, None, :])
    assert torch.equal(only_f.view(B * T, HW, Cc), h2.view(B * T, HW, Cc) + femb[:, None, :])


def test_groupnorm_channels_with_large_mean():
    """Channels with |mean| >> std (mean 50, std 0.5), as a trained residual stream can have: the temporal GroupNorm
    kernels take the variance about the mean (two-pass, like the reference), and the bf16 GEMM epilogue's single-pass
    fixed-point sums must still give a normalised output within the bf16-mode tolerance."""
    o = ops()
    # temporal kernels (register-resident and general)
    for (B, T, HW, Cc) in [(2, 20, 16, 384), (2, 6, 16, 192)]:
        x = rnd(B, T, HW, Cc, seed=1) * 0.5 + 50.0 * torch.sign(rnd(32, seed=9)).repeat_interleave(Cc // 32)   # per GROUP
        gamma, beta = rnd(Cc, seed=2), rnd(Cc, seed=3)
        xr = x.permute(0, 2, 3, 1).reshape(B * HW, Cc, T)
        ref = F.group_norm(xr.double(), 32, gamma.double(), beta.double(), eps=1e-5).view(B, HW, Cc, T).permute(0, 3, 1, 2)
        out, outa = torch.empty_like(x), torch.empty_like(x, dtype=torch.bfloat16)
        o.gn_temporal(x, B, T, HW, Cc, gamma, beta, out, outa)
        torch_err = relerr(F.group_norm(xr, 32, gamma, beta, eps=1e-5).view(B, HW, Cc, T).permute(0, 3, 1, 2), ref)
        assert relerr(out, ref) < max(2e-5, 2 * torch_err)
    # GEMM-epilogue statistics -> GroupNorm apply: a 1x1 conv with identity-like weights and a +-50 bias
    n, H, W, C = 3, 16, 16, 128
    x = (rnd(n, C, H, W, seed=4) * 0.5).bfloat16()
    w = torch.eye(C, device='cuda').bfloat16()
    bias = 50.0 * torch.sign(rnd(32, seed=5)).repeat_interleave(C // 32)
    h = torch.empty(n * H * W, C, device='cuda')
    st = torch.zeros(n, 2, C, device='cuda', dtype=torch.int64)
    o.gemm(nhwc(x.float()).bfloat16(), w, C, n_img=n, H=H, W=W, taps=1, bias=bias, out_f32=h, stats_out=st)
    gamma, beta = rnd(C, seed=6), rnd(C, seed=7)
    ref = F.group_norm(from_nhwc(h, n, H, W).double(), 32, gamma.double(), beta.double(), eps=1e-5)
    out = torch.empty(n * H * W, C, device='cuda', dtype=torch.bfloat16)
    o.gn_apply(h, None, n, H, W, out, stats1=st, gamma=gamma, beta=beta)
    err = relerr(out, nhwc(ref))
    print(f'epilogue-stats GroupNorm with mean/std = 100: max-rel error {err:.3e}')
    assert err < 5e-3          # bf16 output rounding is 4e-3; the statistics must not add to it


def test_cond_mix_and_timestep_embedding():
    from oracle import unet_oracle as U
    o = ops()
    B, Fr, H, W = 2, 3, 8, 8
    x, x0 = rnd(B, Fr, 3, H, W, seed=1), rnd(B, Fr, 3, H, W, seed=2)
    obs = torch.tensor([[1., 0, 0], [1, 1, 0]]).cuda()
    lat = torch.tensor([[0., 1, 0], [0, 0, 1]]).cuda()
    kin = torch.tensor([[0., 0, 0], [0, 0, 0]]).cuda()
    t = torch.tensor([250.0, 999.0]).cuda()
    a = torch.empty(B * Fr * H * W, 64, device='cuda')
    tf = torch.empty(B * Fr, device='cuda')
    am = torch.empty(B * Fr, device='cuda')
    o.cond_mix(x, x0, obs, lat, kin, t, B, Fr, H, W, a, tf, am)
    m = lambda v: v.view(B, Fr, 1, 1, 1)
    any_ = (obs + lat + kin).clamp(max=1)
    ones = torch.ones_like(x[:, :, :1])
    xin = torch.cat([x * m(lat) + x0 * m(obs) + x * (1 - m(any_)), ones * m(obs), ones * m(kin)], 2).view(B * Fr, 5, H, W)
    cols = F.unfold(xin, 3, padding=1).view(B * Fr, 5, 9, H * W).permute(0, 3, 2, 1).reshape(-1, 45)
    assert torch.equal(a[:, :45], cols)
    assert float(a[:, 45:].abs().max()) == 0
    assert torch.equal(tf.view(B, Fr), t.view(B, 1) * (1 - obs))
    assert torch.equal(am.view(B, Fr), any_)
    # cond_emb_type 'duplicate' (6 channels) and 't=0' (x unchanged): unet.py:1014-1019
    o.cond_mix(x, x0, obs, lat, kin, t, B, Fr, H, W, a, tf, am, mode=1)
    xin = torch.cat([x * m(lat) + x * (1 - m(any_)), x0 * m(obs)], 2).view(B * Fr, 6, H, W)
    cols = F.unfold(xin, 3, padding=1).view(B * Fr, 6, 9, H * W).permute(0, 3, 2, 1).reshape(-1, 54)
    assert torch.equal(a[:, :54], cols) and float(a[:, 54:].abs().max()) == 0
    assert torch.equal(tf.view(B, Fr), t.view(B, 1).expand(B, Fr))
    o.cond_mix(x, x0, obs, lat, kin, t, B, Fr, H, W, a, tf, am, mode=2)
    cols = F.unfold(x.view(B * Fr, 3, H, W), 3, padding=1).view(B * Fr, 3, 9, H * W).permute(0, 3, 2, 1).reshape(-1, 27)
    assert torch.equal(a[:, :27], cols) and float(a[:, 27:].abs().max()) == 0
    assert torch.equal(am.view(B, Fr), any_)
    tt = torch.tensor([0.0, 1.0, 17.0, 503.25, 999.0]).cuda()
    emb = torch.empty(5, 128, device='cuda')
    o.timestep_embedding(tt, 128, emb)
    assert float((emb.cpu() - U.sinusoid(tt.cpu(), 128)).abs().max()) < 2e-4   # sin/cos of args up to 1e3
    fi = torch.tensor([0.0, 3.0, 29.0, -4.5, 12.5]).cuda()                     # frame indices, period 10 T
    o.timestep_embedding(fi, 128, emb, max_period=300)
    assert float((emb.cpu() - U.sinusoid(fi.cpu(), 128, max_period=300)).abs().max()) < 1e-5


def _attn_ref(qkv, heads, mask=None, R=None, pad_interact=True):
    """qkv: (B, D, L, 3C) -> (B, D, L, C) following unet.py:477-536."""
    B, Dd, L, C3 = qkv.shape
    Cc = C3 // 3
    hd = Cc // heads
    q, k, v = (qkv.view(B, Dd, L, 3, heads, hd)[:, :, :, i].permute(0, 1, 3, 2, 4) for i in range(3))
    scale = hd ** -0.5
    q = q * scale
    att = q @ k.transpose(-1, -2)
    if R is not None:
        rq, rk, rv = (r.view(B, L, L, heads, hd) for r in R)
        att = att + torch.einsum('bdhtf,btshf->bdhts', q, rk)
        att = att + torch.einsum('bdhtf,btshf->bdhts', k * scale, rq).transpose(-1, -2)
    if mask is not None:
        allowed = mask[:, None, :] * mask[:, :, None]
        if pad_interact:
            allowed = allowed + (1 - mask[:, None, :]) * (1 - mask[:, :, None])
        else:
            i = torch.arange(L)
            allowed[:, i, i] = 1
        att = att.masked_fill((allowed == 0).view(B, 1, 1, L, L), float('-inf'))
    w = torch.softmax(att, -1)
    out = w @ v
    if R is not None:
        out = out + torch.einsum('bdhts,btshf->bdhtf', w, rv)
    return out.permute(0, 1, 3, 2, 4).reshape(B, Dd, L, Cc)


@pytest.mark.parametrize('T,hd,pad', [(20, 96, True), (10, 32, False), (7, 128, True)])  # noqa
def test_attention_temporal(T, hd, pad):
    o = ops()
    B, HW, heads = 2, 19, 4
    Cc = heads * hd
    qkv = rnd(B, T, HW, 3 * Cc, seed=1)
    R = [rnd(B * T * T, Cc, seed=2 + i, scale=0.5) for i in range(3)]
    mask = torch.ones(B, T).cuda()
    mask[0, T - 2:] = 0
    ref = _attn_ref(qkv.permute(0, 2, 1, 3), heads, mask, R, pad).permute(0, 2, 1, 3)   # back to (B,T,HW,C)
    out = torch.empty(B, T, HW, Cc, device='cuda')
    o.attn_temporal(qkv, R[0], R[1], R[2], mask, pad, B, T, HW, heads, hd, out)
    assert relerr(out, ref) < 2e-5


@pytest.mark.parametrize('T,hd,HW,pad', [(20, 96, 256, True), (10, 32, 64, False), (7, 128, 128, True),
                                          (20, 128, 64, True)])
def test_attention_temporal_tensor_core_path(T, hd, HW, pad):
    """RPE terms as grouped tcgen05 GEMMs + mma.sync attention core, against the fp32 einsum restatement."""
    o = ops()
    B, heads = 3, 4
    Cc = heads * hd
    M = B * T * HW
    qkv = rnd(B, T, HW, 3 * Cc, seed=1).bfloat16()
    R = [rnd(B * T * T, Cc, seed=2 + i, scale=0.5) for i in range(3)]
    mask = torch.ones(B, T).cuda()
    mask[0, T - 2:] = 0
    # reference on the same bf16-rounded q, k, v and bf16-rounded R tables
    Rb = [r.bfloat16().float() for r in R]
    Rb[0] = (R[0] * hd ** -0.5).bfloat16().float() * hd ** 0.5           # bq is rounded after the scale
    ref = _attn_ref(qkv.float().permute(0, 2, 1, 3), heads, mask, Rb, pad).permute(0, 2, 1, 3).reshape(M, Cc)
    gpt, tpg = (1, HW // 128) if HW >= 128 else (128 // HW, 1)
    SW, ntg = 128 * gpt, (B * T + gpt - 1) // gpt
    bq = torch.empty(ntg * SW, Cc, device='cuda', dtype=torch.bfloat16)
    bk = torch.empty_like(bq)
    bv = torch.empty(ntg * Cc, SW, device='cuda', dtype=torch.bfloat16)
    o.rpe_expand(R[0], R[1], R[2], B, T, heads, hd, gpt, bq, bk, bv)      # (bias=None: tables already complete)
    q2 = qkv.view(M, 3 * Cc)
    lin = dict(n_img=M, H=1, W=1, taps=1)
    sk, sq = torch.empty(M, SW, device='cuda'), torch.empty(M, SW, device='cuda')
    o.gemm(q2[:, :Cc], bk, SW, out_f32=sk, w_group_tiles=tpg, C1=Cc, **lin)
    o.gemm(q2[:, Cc:2 * Cc], bq, SW, out_f32=sq, w_group_tiles=tpg, C1=Cc, **lin)
    # spot-check the grouped GEMM: Sk[(b,t,d)][(h,s)] = q . Rk[b,t,s,h,:]
    qf = qkv.float().view(B, T, HW, 3, heads, hd)[:, :, :, 0]
    sk_ref = torch.einsum('btdhf,btshf->btdhs', qf, Rb[1].view(B, T, T, heads, hd))
    got = sk.view(B, T, HW, SW)
    for t in range(T):
        sub = torch.stack([got[b, t, :, ((b * T + t) % gpt) * 128:((b * T + t) % gpt) * 128 + heads * T] for b in range(B)])
        assert relerr(sub.reshape(B, HW, heads, T), sk_ref[:, t]) < 2e-5
    pm = torch.zeros(M, SW, device='cuda', dtype=torch.bfloat16)
    pv = torch.empty(M, Cc, device='cuda')
    o.attn_temporal_tc(q2, sk, sq, mask, pad, B, T, HW, heads, hd, gpt, pm, pv)
    att = torch.empty(M, Cc, device='cuda', dtype=torch.bfloat16)
    o.gemm(pm, bv, Cc, residual=pv, out_bf16=att, w_group_tiles=tpg, **lin)
    assert relerr(att, ref) < 1.2e-2        # P and the attention output are rounded to bf16


def _rpe_fragment_major(Rb, B, T, heads, hd, which):
    """Expected vdm_rpe_pack output (see vdm.h): 16-byte vectors = the four A registers of an mma lane.  Rb: [B*T*T][C]
    bf16-rounded (R + bias) as fp32.  which 'qk': [G][heads][2][hd/32][2][32][4][2]; 'v': [G][heads][hd/16][2][32][4][2]."""
    G = B * T
    R5 = Rb.view(G, T, heads, hd).cpu()
    ar = torch.arange
    if which == 'qk':
        mt, p, kk2, lane, r, e = torch.meshgrid(ar(2), ar(hd // 32), ar(2), ar(32), ar(4), ar(2), indexing='ij')
        row = 16 * mt + lane // 4 + 8 * (r & 1)                                  # second frame index j
        con = 32 * p + 8 * (lane % 4) + 4 * kk2 + 2 * (r >> 1) + e              # channel f
        vals = R5[:, row.clamp(max=T - 1), :, con]                              # index dims first, then (G, heads)
        vals = vals * (row < T)[..., None, None]
    else:
        mt, kk2, lane, r, e = torch.meshgrid(ar(hd // 16), ar(2), ar(32), ar(4), ar(2), indexing='ij')
        row = 16 * mt + lane // 4 + 8 * (r & 1)                                  # channel f
        con = 8 * (lane % 4) + 4 * kk2 + 2 * (r >> 1) + e                       # second frame index s
        vals = R5[:, con.clamp(max=T - 1), :, row]
        vals = vals * (con < T)[..., None, None]
    nd = vals.dim()
    return vals.permute(nd - 2, nd - 1, *range(nd - 2)).contiguous()


@pytest.mark.parametrize('T,hd,HW,pad,pt', [(20, 96, 256, True, 8), (20, 96, 256, True, 16), (10, 32, 64, False, 8),
                                             (7, 128, 128, True, 8), (20, 128, 64, True, 8), (32, 64, 24, False, 8),
                                             (27, 96, 16, True, 8), (1, 96, 8, True, 8), (17, 96, 48, False, 16)])
def test_attention_temporal_fused(T, hd, HW, pad, pt):
    """The whole temporal attention of a block in one kernel (RPE score terms, q.k^T, mask, softmax, P.V, attn.R_v),
    against the fp32 einsum restatement on the same bf16-rounded q, k, v and bf16-rounded (R + bias) tables."""
    o = ops()
    B, heads = 3, 4
    Cc = heads * hd
    M = B * T * HW
    qkv = rnd(B, T, HW, 3 * Cc, seed=1).bfloat16()
    R = [rnd(B * T * T, Cc, seed=2 + i, scale=0.5) for i in range(3)]
    bias = rnd(3, Cc, seed=9, scale=0.2)
    mask = torch.ones(B, T).cuda()
    mask[0, max(T - 2, 1):] = 0
    Rb = [(R[i] + bias[i]).bfloat16().float() for i in range(3)]
    ref = _attn_ref(qkv.float().permute(0, 2, 1, 3), heads, mask, Rb, pad).permute(0, 2, 1, 3).reshape(M, Cc)
    TP = 24 if T <= 24 else 32
    nan = float('nan')
    per = heads * hd * 32                         # elements per (b, t) group and table
    rq = torch.full((B * T, per), nan, device='cuda', dtype=torch.bfloat16)
    rk, rv = torch.full_like(rq, nan), torch.full_like(rq, nan)
    o.rpe_pack(R[0], R[1], R[2], B, T, heads, hd, rq, rk, rv, bias=bias)
    assert torch.equal(rq.float().cpu().view(-1), _rpe_fragment_major(Rb[0], B, T, heads, hd, 'qk').view(-1))
    assert torch.equal(rk.float().cpu().view(-1), _rpe_fragment_major(Rb[1], B, T, heads, hd, 'qk').view(-1))
    assert torch.equal(rv.float().cpu().view(-1), _rpe_fragment_major(Rb[2], B, T, heads, hd, 'v').view(-1))
    att = torch.full((M, Cc), nan, device='cuda', dtype=torch.bfloat16)
    o.attn_temporal_fused(qkv.view(M, 3 * Cc), rq, rk, rv, mask, pad, B, T, HW, heads, hd, TP, att, pixels_per_cta=pt)
    assert relerr(att, ref) < 1.2e-2        # P and the attention output are rounded to bf16
    # a micro-batch: videos 1.. of the same tables through offset views
    att2 = torch.full((M, Cc), nan, device='cuda', dtype=torch.bfloat16)
    o.attn_temporal_fused(qkv.view(M, 3 * Cc)[T * HW:], rq[T:], rk[T:], rv[T:], mask[1:], pad, B - 1, T, HW, heads, hd, TP,
                          att2[T * HW:], pixels_per_cta=pt)
    assert torch.equal(att2[T * HW:], att[T * HW:])


def test_rpe_pack_batched_over_blocks_and_unsupported_shapes():
    o = ops()
    B, T, heads, hd, nb = 2, 20, 4, 32, 3
    Cc, rows = heads * hd, 2 * 20 * 20
    Rall = rnd(nb * 3 * rows, Cc, seed=3)
    bias = rnd(nb, 3, Cc, seed=4)
    per = B * T * heads * hd * 32
    rqk = torch.zeros(nb, 2, per, device='cuda', dtype=torch.bfloat16)
    rvp = torch.zeros(nb, per, device='cuda', dtype=torch.bfloat16)
    o.rpe_pack(Rall[:rows], Rall[rows:2 * rows], Rall[2 * rows:3 * rows], B, T, heads, hd, rqk[0, 0], rqk[0, 1], rvp,
               bias=bias, n_blocks=nb, r_block_stride=3 * rows * Cc, qk_block_stride=2 * per)
    for i in range(nb):
        blk = Rall[i * 3 * rows:(i + 1) * 3 * rows].view(3, rows, Cc)
        q1, k1 = torch.empty(per, device='cuda', dtype=torch.bfloat16), torch.empty(per, device='cuda', dtype=torch.bfloat16)
        v1 = torch.empty_like(rvp[0])
        o.rpe_pack(blk[0], blk[1], blk[2], B, T, heads, hd, q1, k1, v1, bias=bias[i])
        assert torch.equal(rqk[i, 0], q1) and torch.equal(rqk[i, 1], k1) and torch.equal(rvp[i], v1)
    out = torch.empty(B * T * 8, Cc, device='cuda', dtype=torch.bfloat16)
    qkv = torch.zeros(B * T * 8, 3 * Cc, device='cuda', dtype=torch.bfloat16)
    mask = torch.ones(B, T).cuda()
    with pytest.raises(RuntimeError, match='pixels per CTA'):
        o.attn_temporal_fused(qkv, rqk[0, 0], rqk[0, 1], rvp[0], mask, True, B, T, 8, heads, hd, 24, out, pixels_per_cta=16)
    with pytest.raises(RuntimeError, match='t_pad'):
        o.attn_temporal_fused(qkv, rqk[0, 0], rqk[0, 1], rvp[0], mask, True, B, 33, 8, heads, hd, 32, out)
    # shapes beyond the 227 KB of shared memory are reported, the model falls back to the three-launch path for them
    assert o.attn_temporal_fused_smem(20, 96, 24, 8) == 3 * 20 * (8 * 192 + 16) + 8 * 1952 + 32
    assert o.attn_temporal_fused_smem(32, 128, 32, 8) > 227 * 1024 and o.attn_temporal_fused_smem(20, 128, 24, 16) == -1


@pytest.mark.parametrize('M,C,SW,tpg', [(640, 128, 128, 1), (1536, 64, 256, 1), (1024, 192, 128, 2)])
def test_gemm_problem_batch_equals_separate_launches(M, C, SW, tpg):
    """n_prob = 2: q -> Sk and k -> Sq (column blocks of one qkv matrix, stacked grouped weights, stacked outputs) in
    one launch, bit-identical to the two separate grouped GEMMs."""
    o = ops()
    qkv = rnd(M, 3 * C, seed=1).bfloat16()
    groups = (M // 128 + tpg - 1) // tpg
    rows = groups * SW
    bkq = rnd(2, rows, C, seed=2, scale=C ** -0.5).bfloat16()
    lin = dict(n_img=M, H=1, W=1, taps=1)
    sk, sq = torch.empty(M, SW, device='cuda'), torch.empty(M, SW, device='cuda')
    o.gemm(qkv[:, :C], bkq[0], SW, out_f32=sk, w_group_tiles=tpg, C1=C, **lin)
    o.gemm(qkv[:, C:2 * C], bkq[1], SW, out_f32=sq, w_group_tiles=tpg, C1=C, **lin)
    both = torch.full((2, M, SW), float('nan'), device='cuda')
    o.gemm(qkv[:, :C], bkq.view(2 * rows, C), SW, out_f32=both, w_group_tiles=tpg, C1=C, n_prob=2, prob_a_cols=C,
           prob_w_rows=rows, prob_out_stride=M * SW, **lin)
    assert torch.equal(both[0], sk) and torch.equal(both[1], sq)
    ref = torch.einsum('gmk,gnk->gmn', qkv[:, :C].float().view(groups, -1, C) if tpg == 1 and M // 128 == groups else
                       qkv[:, :C].float().view(groups, -1, C), bkq[0].float().view(groups, SW, C)).reshape(M, SW)
    assert relerr(sk, ref) < 2e-5


@pytest.mark.parametrize('L,hd', [(256, 96), (64, 128), (256, 32)])
def test_attention_spatial_f32(L, hd):
    o = ops()
    n, heads = 3, 4
    Cc = heads * hd
    qkv = rnd(n, L, 3 * Cc, seed=1)
    ref = _attn_ref(qkv.view(n, 1, L, 3 * Cc), heads).view(n, L, Cc)
    out = torch.empty(n, L, Cc, device='cuda')
    o.attn_spatial(qkv, n, L, heads, hd, out)
    assert relerr(out, ref) < 2e-5


@pytest.mark.parametrize('L,hd', [(256, 96), (64, 128), (256, 32), (100, 64)])
def test_attention_spatial_tensor_core(L, hd):
    o = ops()
    n, heads = 3, 4
    Cc = heads * hd
    qkv = rnd(n, L, 3 * Cc, seed=1).bfloat16()
    ref = _attn_ref(qkv.float().view(n, 1, L, 3 * Cc), heads).view(n, L, Cc)
    for odt in (torch.float32, torch.bfloat16):
        out = torch.empty(n, L, Cc, device='cuda', dtype=odt)
        o.attn_spatial(qkv, n, L, heads, hd, out)
        assert relerr(out, ref) < (4e-3 if odt == torch.float32 else 8e-3)   # P is rounded to bf16 for the PV product


@pytest.mark.parametrize('n,L,hd', [(160, 256, 96), (161, 64, 128), (5, 128, 64), (37, 256, 128), (9, 64, 32),
                                    (150, 128, 96)])
def test_attention_spatial_tcgen05(n, L, hd):
    """The tcgen05 / TMEM kernel (bf16 in, bf16 out, L in {64, 128, 256}): several work items per CTA, two images per
    tile at L = 64 incl. an odd image count, both softmax groups, peaky and flat rows."""
    o = ops()
    heads = 4
    Cc = heads * hd
    qkv = (rnd(n, L, 3 * Cc, seed=3) * 1.7).bfloat16()
    qkv[0, :, :Cc] *= 4.0          # one image with sharply peaked rows
    qkv[-1, :, :Cc] = 0            # and one with exactly uniform attention
    ref = _attn_ref(qkv.float().view(n, 1, L, 3 * Cc), heads).view(n, L, Cc)
    out = torch.full((n, L, Cc), float('nan'), device='cuda', dtype=torch.bfloat16)
    o.attn_spatial(qkv, n, L, heads, hd, out)
    assert torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 8e-3
    per_img = (out.float() - ref).flatten(1).abs().amax(1) / ref.flatten(1).abs().amax(1)
    assert float(per_img.max()) < 2e-2, per_img


def test_rpe_hidden():
    o = ops()
    B, T, Cc = 2, 6, 64
    e_t = rnd(B * T, 3 * Cc, seed=1)
    fi = torch.tensor([[0, 1, 2, 10, 11, 30], [5, 3, 3, 100, 7, 8]]).cuda()
    wd, bd = rnd(3, Cc, 3, seed=2), rnd(3, Cc, seed=3)
    out = torch.empty(3, B * T * T, Cc, device='cuda')
    o.rpe_hidden(e_t, fi, wd, bd, B, T, Cc, out)
    d = (fi[:, :, None] - fi[:, None, :]).float()
    feats = torch.stack([torch.log1p(d.clamp(min=0)), torch.log1p((-d).clamp(min=0)), (d == 0).float()], -1)
    for net in range(3):
        ed = feats @ wd[net].t() + bd[net]
        ref = F.silu(e_t.view(B, T, 1, 3, Cc)[:, :, :, net] + ed).reshape(B * T * T, Cc)
        assert relerr(out[net], ref) < 1e-5


def test_rpe_tables_batched_over_blocks_equal_per_block_calls():
    """n_blocks > 1 (the model's one-launch-per-(C, HW)-group path) is bit-identical to per-block launches."""
    o = ops()
    B, T, heads, hd, nb, HW = 2, 8, 4, 32, 3, 64
    Cc, rows = heads * hd, B * T * T
    width = 3 * Cc * nb + 40
    e_t = rnd(B * T, width, seed=1)
    offs = [40, 40 + 6 * Cc, 40 + 3 * Cc]                      # deliberately not in order
    fi = torch.randint(0, 50, (B, T), generator=torch.Generator().manual_seed(0)).cuda()
    wd, bd = rnd(nb * 3, Cc, 3, seed=2), rnd(nb * 3, Cc, seed=3)
    hid = torch.empty(nb, 3, rows, Cc, device='cuda', dtype=torch.bfloat16)
    o.rpe_hidden(e_t, fi, wd, bd, B, T, Cc, hid, et_offsets=torch.tensor(offs, dtype=torch.int32).cuda(), n_blocks=nb)
    for i in range(nb):
        one = torch.empty(3, rows, Cc, device='cuda', dtype=torch.bfloat16)
        o.rpe_hidden(e_t[:, offs[i]:offs[i] + 3 * Cc], fi, wd[3 * i:3 * i + 3], bd[3 * i:3 * i + 3], B, T, Cc, one)
        assert torch.equal(one, hid[i])
    R = rnd(nb, 3, rows, Cc, seed=4, scale=0.5)
    bias = rnd(nb, 3, Cc, seed=5)
    gpt = 128 // HW
    SW, ntg = 128 * gpt, (B * T + gpt - 1) // gpt
    bq = torch.empty(nb, ntg * SW, Cc, device='cuda', dtype=torch.bfloat16)
    bk = torch.empty_like(bq)
    bv = torch.empty(nb, ntg * Cc, SW, device='cuda', dtype=torch.bfloat16)
    o.rpe_expand(R[0, 0], R[0, 1], R[0, 2], B, T, heads, hd, gpt, bq, bk, bv, bias=bias, n_blocks=nb,
                 r_block_stride=3 * rows * Cc)
    for i in range(nb):
        q1, k1, v1 = torch.empty_like(bq[0]), torch.empty_like(bk[0]), torch.empty_like(bv[0])
        o.rpe_expand(R[i, 0], R[i, 1], R[i, 2], B, T, heads, hd, gpt, q1, k1, v1, bias=bias[i])
        assert torch.equal(q1, bq[i]) and torch.equal(k1, bk[i]) and torch.equal(v1, bv[i])
    # zero_fill=False rewrites only the live (block-diagonal) entries of buffers that were zeroed once
    zq, zk, zv = torch.zeros_like(bq), torch.zeros_like(bk), torch.zeros_like(bv)
    for _ in range(2):
        o.rpe_expand(R[0, 0], R[0, 1], R[0, 2], B, T, heads, hd, gpt, zq, zk, zv, bias=bias, n_blocks=nb,
                     r_block_stride=3 * rows * Cc, zero_fill=False)
    assert torch.equal(zq, bq) and torch.equal(zk, bk) and torch.equal(zv, bv)


@pytest.mark.parametrize('alpha,beta,gamma', [(3, 7, 20.5), (30, 30, 30), (2, 5, 9.5), (1, 12, 40.25)])
def test_rpe_lookup_matches_oracle_bucket_function(alpha, beta, gamma):
    """use_rpe_net=False: table[bucket(distance)] with the reference's piecewise bucket function (oracle restatement
    of unet.py:326-347; gamma is kept off the integer grid, where the reference itself is 1-ulp fragile)."""
    from oracle import unet_oracle as U
    o = ops()
    B, T, Cc = 2, 9, 64
    fi = torch.tensor([[0, 1, 2, 3, 5, 8, 13, 21, 29], [29, 4, 4, 0, 17, 18, 9, 25, 2]])
    tables = rnd(3, 2 * beta + 1, Cc, seed=3)
    out = torch.empty(3, B * T * T, Cc, device='cuda')
    o.rpe_lookup(tables, fi.cuda(), B, T, Cc, alpha, beta, gamma, out)
    dist = fi[:, :, None] - fi[:, None, :]
    ids = U.rpe_bucket_ids(dist, alpha, beta, gamma)
    for net in range(3):
        ref = tables[net].cpu()[ids].reshape(B * T * T, Cc)
        assert torch.equal(out[net].cpu(), ref)


@pytest.mark.parametrize('respacing', ['', 'ddim10'])
def test_sampler_kernels_match_oracle(respacing):
    from video_diffusion_b200.gaussian_diffusion import device_tables
    o = ops()
    s = D.Schedule(1000, 'linear', respacing)
    tab = device_tables(s_like=s).cuda()
    shape = (3, 4, 3, 8, 8)
    x, eps, z, x0 = (rnd(*shape, seed=i) for i in range(4))
    x0 = x0.clamp(-1, 1)
    n = s.num_timesteps
    for tl in ([0, 1, n - 1], [n // 2, 0, 3]):
        t = torch.tensor(tl).cuda()
        ref = D.p_sample(s, eps.cpu(), x.cpu(), t.cpu(), z.cpu())
        pred = torch.empty_like(x)
        got = o.sampler_step(0, x, eps, z, t, tab, pred_xstart=pred)
        assert float((got.cpu() - ref['sample']).abs().max()) < 2e-6
        assert float((pred.cpu() - ref['pred_xstart']).abs().max()) < 2e-6
        for eta in (0.0, 0.7):
            ref = D.ddim_sample(s, eps.cpu(), x.cpu(), t.cpu(), z.cpu(), eta=eta)
            got = o.sampler_step(1, x, eps, z, t, tab, eta=eta)
            assert float((got.cpu() - ref['sample']).abs().max()) < 4e-6
        assert float((o.q_sample(x0, z, t, tab).cpu() - D.q_sample(s, x0.cpu(), t.cpu(), z.cpu())).abs().max()) < 1e-6
        lat = torch.tensor([[0., 1, 1, 0], [1, 1, 1, 1], [0, 0, 0, 1]]).cuda()
        xt = D.q_sample(s, x0.cpu(), t.cpu(), z.cpu())
        ref = D.vb_terms(s, eps.cpu(), x0.cpu(), xt, t.cpu(), lat.cpu().view(3, 4, 1, 1, 1))
        acc = torch.zeros(3, 3, device='cuda', dtype=torch.float64)
        o.vb_terms(x0, xt.cuda(), eps, z, t, tab, lat, True, acc)
        np.testing.assert_allclose(acc[:, 0].cpu().numpy(), ref['output'].numpy(), rtol=2e-5, atol=1e-7)
        ref_x = D.mean_flat((ref['pred_xstart'] - x0.cpu()) ** 2, lat.cpu().view(3, 4, 1, 1, 1))
        np.testing.assert_allclose(acc[:, 1].cpu().numpy(), ref_x.numpy(), rtol=2e-5)
    acc = torch.zeros(3, device='cuda', dtype=torch.float64)
    o.prior_bpd(x0, tab, lat, acc)
    np.testing.assert_allclose(acc.cpu().numpy(), D.prior_bpd(s, x0.cpu(), lat.cpu().view(3, 4, 1, 1, 1)).numpy(), rtol=2e-5,
                               atol=2e-7)   # -1 - lv + e^lv + m^2 cancels to ~1e-5 in fp32
