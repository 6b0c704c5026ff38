"""Tensor-level wrappers over the C ABI (one Python function per libvdm entry point).

Every function launches asynchronously on torch's current CUDA stream and returns the
output tensors it was given / allocated.  Nothing here computes on the host.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F16, F32, GemmArgs, GnApplyArgs, check, dt, ptr, stream

# Optional per-launch timing (bench.py / profiling only): when PROFILE is a list, every wrapper
# brackets its launch with CUDA events on the current stream and appends
# (kernel name, start event, end event, algorithmic flops, algorithmic bytes).
PROFILE = None
# Timeline mode (profiles/step_timeline.py): with PROFILE_STREAMS a list as well, the model keeps its micro-batch and
# side streams while profiling, and every record's stream handle is appended here (same index as in PROFILE).
PROFILE_STREAMS = None


def _timed(name, fn, flops=0.0, nbytes=0.0, meta=''):
    if PROFILE is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    PROFILE.append((name, e0, e1, flops, nbytes, meta))
    if PROFILE_STREAMS is not None:
        PROFILE_STREAMS.append(torch.cuda.current_stream().cuda_stream)


def _nbytes(*tensors):
    return float(sum(t.numel() * t.element_size() for t in tensors if t is not None))


def _gemm_args(a1, w, N, *, n_img, H, W, taps, a1_mode=0, a2=None, bias=None, rowbias=None, residual=None,
               out_f32=None, out_bf16=None, out_nchw=False, out_silu=None, C1=None, C2=0, stats_out=None,
               w_group_tiles=0, n_prob=1, prob_a_cols=0, prob_w_rows=0, prob_out_stride=0, a1_coef=None, a1_act=False,
               a2b=None, img_done=None):
    g = GemmArgs()
    g.dtype = dt(a1.dtype)
    if a1_coef is not None and out_nchw and a1.dtype == torch.float16:
        # output head with the GroupNorm-apply fused into its staging: a1 = the raw fp16 stream, weights bf16
        g.dtype, g.a1_raw_dtype = BF16, F16
    g.taps, g.a1_mode = taps, a1_mode
    g.n_img, g.H, g.W = n_img, H, W
    g.C1 = a1.shape[-1] if C1 is None else C1
    g.C2 = C2 if a2 is None else a2.shape[-1]
    g.N = N
    if a1.dim() == 2 and not a1.is_contiguous() and a1.stride(1) == 1:      # column block of a wider matrix
        g.a1, g.lda1 = a1.data_ptr(), a1.stride(0)
    else:
        g.a1, g.lda1 = ptr(a1), 0
    g.a2, g.w = ptr(a2), ptr(w)
    # second operand range: one or two tensors (channel concat), bf16 -- or fp16 (the stream itself) with fp16 weight
    # columns for that range
    g.a2b, g.C2b = ptr(a2b), (0 if a2b is None else a2b.shape[-1])
    if a2 is not None:
        if a2.dtype not in (torch.bfloat16, torch.float16, torch.float32) or (a2b is not None and a2b.dtype != a2.dtype):
            raise TypeError('vdm_gemm: a2 / a2b must share a 16-bit dtype')
        g.a2_dtype = dt(a2.dtype) if a2.dtype != torch.float32 else 0
    g.w_group_tiles = w_group_tiles
    g.n_prob, g.prob_a_cols, g.prob_w_rows, g.prob_out_stride = n_prob, prob_a_cols, prob_w_rows, prob_out_stride
    g.bias = ptr(bias)
    g.rowbias = None if rowbias is None else rowbias.data_ptr()
    g.ld_rowbias = rowbias.stride(0) if rowbias is not None else 0
    g.residual = ptr(residual)
    g.ld_res = residual.shape[-1] if residual is not None else 0
    g.out_f32, g.out_bf16 = ptr(out_f32), ptr(out_bf16)
    g.ld_out = out_f32.shape[-1] if (out_f32 is not None and not out_nchw) else (
        out_silu.shape[-1] if out_silu is not None else 0)
    g.ld_out_bf16 = out_bf16.shape[-1] if out_bf16 is not None else 0
    g.out_nchw = int(out_nchw)
    g.out_silu_f32 = ptr(out_silu)
    g.stats_out = ptr(stats_out)
    g.a1_coef, g.a1_act = ptr(a1_coef, torch.float32), int(a1_act)
    g.img_done = None if img_done is None else _counter_ptr(img_done, n_img)
    # `out_f32` / `residual` are the model's residual stream: fp32, or fp16 in the bf16 model (io_dtype)
    io = {t.dtype for t in (out_f32, residual) if t is not None}
    if len(io) > 1 or not io <= {torch.float32, torch.float16}:
        raise TypeError(f'vdm_gemm: out_f32 / residual must both be fp32 or both fp16, got {io}')
    g.io_dtype = dt(io.pop()) if io else F32
    return g


def _counter_ptr(t, n_img):
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous() and t.numel() >= n_img):
        raise TypeError('per-image completion counters: a contiguous int32 CUDA tensor of n_img entries')
    return t.data_ptr()


def gemm(a1, w, N, **kw):
    """Implicit-GEMM conv / linear (see include/vdm.h: vdm_gemm).  `a1_coef` (from gn_coef) fuses the GroupNorm-apply
    (+ SiLU with a1_act) of the A operand into the kernel: a1 is then the RAW bf16 activation."""
    lib = _lib.load()
    g = _gemm_args(a1, w, N, **kw)
    M = g.n_img * g.H * g.W
    K = g.taps * g.C1 + g.C2 + g.C2b
    name = ('gemm_tc' if g.dtype in (BF16, F16) else 'gemm_simt') + ('_conv3x3' if g.taps == 9 else '_linear')
    _timed(name, lambda: check(lib.vdm_gemm(C.byref(g), stream()), 'vdm_gemm'), flops=2.0 * M * N * K * g.n_prob,
           meta=f'M={M} N={N} K={K} HxW={g.H}x{g.W} mode={g.a1_mode} res={int(bool(g.residual))} '
                f'stats={int(bool(g.stats_out))}' + (' norm=fused' if g.a1_coef else '') +
                (f' x{g.n_prob}' if g.n_prob > 1 else ''))


def gemm_fused_norm_supported(a1, w, N, **kw):
    """Would this call (with a1_coef) run on a kernel that has the fused-normalisation stage?  Same dispatch code as
    vdm_gemm, no launch."""
    g = _gemm_args(a1, w, N, **kw)
    return bool(_lib.load().vdm_gemm_fused_norm_supported(C.byref(g)))


def gemm_img_done_supported(a1, w, N, **kw):
    """Would this call run on a kernel that maintains per-image completion counters (`img_done`)?  No launch."""
    g = _gemm_args(a1, w, N, **kw)
    return bool(_lib.load().vdm_gemm_img_done_supported(C.byref(g)))


def gn_coef(stats1, stats2, n_img, HW, gamma, beta, coef, scale_shift=None):
    """Per-(image, channel) (a, b) of GroupNorm32 (+ scale/shift) into coef [n_img][C1+C2][2] (see vdm.h: vdm_gn_coef)."""
    kind = lambda s: _lib.F64 if (s is not None and s.dtype == torch.float64) else _lib.I64
    C1, C2 = stats1.shape[-1], (0 if stats2 is None else stats2.shape[-1])
    _timed('gn_coef', lambda: check(_lib.load().vdm_gn_coef(
        ptr(stats1), kind(stats1), C1, ptr(stats2), kind(stats2), C2, n_img, HW, ptr(gamma, torch.float32),
        ptr(beta, torch.float32), None if scale_shift is None else scale_shift.data_ptr(),
        0 if scale_shift is None else scale_shift.stride(0), ptr(coef, torch.float32), stream()), 'vdm_gn_coef'),
           nbytes=_nbytes(coef))


def gn_stats(src, n_img, HW, stats):
    """Per-(image, channel) sum / sum of squares of src [n_img*HW][C] into stats [n_img][2][C] (float64, zeroed)."""
    lib = _lib.load()
    _timed('gn_stats', lambda: check(lib.vdm_gn_stats_t(ptr(src), dt(src.dtype), src.shape[-1], n_img, HW,
                                                        ptr(stats, torch.float64), stream()), 'vdm_gn_stats'),
           nbytes=_nbytes(src))


def gn_apply(src1, src2, n_img, H, W, out, *, stats1=None, stats2=None, gamma=None, beta=None, scale_shift=None,
             silu=False, out_mode=0, copy=None, out_raw=None, wait_done=None):
    lib = _lib.load()
    a = GnApplyArgs()
    a.src1, a.C1 = ptr(src1), src1.shape[-1]
    a.src2, a.C2 = ptr(src2), (0 if src2 is None else src2.shape[-1])
    a.n_img, a.H, a.W = n_img, H, W
    a.stats1, a.stats2 = ptr(stats1), ptr(stats2)
    a.stats_dtype = _lib.F64 if (stats1 is not None and stats1.dtype == torch.float64) else _lib.I64
    a.stats2_dtype = _lib.F64 if (stats2 is not None and stats2.dtype == torch.float64) else _lib.I64
    a.gamma, a.beta = ptr(gamma), ptr(beta)
    a.scale_shift = None if scale_shift is None else scale_shift.data_ptr()
    a.ld_ss = 0 if scale_shift is None else scale_shift.stride(0)
    # `out` may be None when only the `copy` output is wanted (the fp16 stream copy that feeds the qkv projection)
    a.silu, a.out_mode, a.out_dtype = int(silu), out_mode, (BF16 if out is None else dt(out.dtype))
    a.out, a.out_raw, a.out_f32_copy = ptr(out), ptr(out_raw), ptr(copy)
    a.src1_dtype = dt(src1.dtype)
    a.copy_dtype = F32 if copy is None else dt(copy.dtype)
    if wait_done is not None:       # launched right behind the conv that produces src1: normalise image by image
        a.wait_done, a.wait_count = _counter_ptr(wait_done, n_img), H * W * a.C1
    if src2 is not None and src2.dtype != src1.dtype:
        raise TypeError('vdm_gn_apply: the two sources of a concat must share a dtype')
    _timed('gn_apply', lambda: check(lib.vdm_gn_apply(C.byref(a), stream()), 'vdm_gn_apply'),
           nbytes=_nbytes(src1, src2, out, copy, out_raw),
           meta=f'n={n_img} HxW={H}x{W} C={a.C1}+{a.C2} in={src1.dtype} norm={int(stats1 is not None)} '
                f'silu={int(silu)} mode={out_mode} raw={int(out_raw is not None)} copy={int(copy is not None)}')


def gn_temporal(x, B, T, HW, Cc, gamma, beta, out_f32, out_a):
    lib = _lib.load()
    if out_f32 is not None and out_f32.dtype != x.dtype:
        raise TypeError('vdm_gn_temporal: the normalised residual copy has the dtype of x')
    out_dt = BF16 if out_a is None else dt(out_a.dtype)        # out_a None: the normalised copy alone
    _timed('gn_temporal', lambda: check(lib.vdm_gn_temporal_t(ptr(x), dt(x.dtype), B, T, HW, Cc, ptr(gamma), ptr(beta),
                                                              ptr(out_f32), ptr(out_a), out_dt, stream()),
                                        'vdm_gn_temporal'), nbytes=_nbytes(x, out_f32, out_a))


def add_spatial_encoding(h, enc, out, n_img, HW, Cc, frame_emb=None):
    _timed('add_spatial_encoding', lambda: check(_lib.load().vdm_add_spatial_encoding_t(
        ptr(h), dt(h.dtype), ptr(enc, torch.float32), ptr(frame_emb, torch.float32), ptr(out, h.dtype), n_img, HW, Cc,
        stream()), 'vdm_add_spatial_encoding'), nbytes=_nbytes(h, out))


def cond_mix(x, x0, obs, lat, kinda, t, B, F, H, W, a_out, t_frame, attn_mask, mode=0):
    _timed('cond_mix', lambda: check(_lib.load().vdm_cond_mix(
        ptr(x), ptr(x0), ptr(obs), ptr(lat), ptr(kinda), ptr(t), B, F, H, W, mode, ptr(a_out), dt(a_out.dtype), ptr(t_frame),
        ptr(attn_mask), stream()), 'vdm_cond_mix'), nbytes=_nbytes(x, x0, a_out))


def map_timesteps(t, tmap, scale):
    """respace.py:113-119 in one launch: float32 (timestep_map[clamp(t)] * scale)."""
    out = torch.empty(t.shape, device=t.device, dtype=torch.float32)
    _timed('map_timesteps', lambda: check(_lib.load().vdm_map_timesteps(
        ptr(t, torch.long), ptr(tmap, torch.long), tmap.numel(), float(scale), ptr(out), t.numel(), stream()),
        'vdm_map_timesteps'))
    return out


def stage_inputs(x, x0, obs, lat, kinda, t, fi, ws):
    """The inputs of a forward into the workspace tensors of `ws`, one launch (vdm_stage_inputs); fi may be None."""
    B, F = ws.B, ws.F
    _timed('stage_inputs', lambda: check(_lib.load().vdm_stage_inputs(
        ptr(x, torch.float32), ptr(x0, torch.float32), ptr(obs, torch.float32), ptr(lat, torch.float32),
        ptr(kinda, torch.float32), ptr(t, torch.float32), ptr(fi, torch.long), B, F, x.numel(), ptr(ws.x), ptr(ws.x0),
        ptr(ws.obs), ptr(ws.lat), ptr(ws.kinda), ptr(ws.t), ptr(ws.fi), stream()), 'vdm_stage_inputs'),
           nbytes=4.0 * x.numel() * 4)


def timestep_embedding(t_frame, dim, out, max_period=10000.0):
    _timed('timestep_embedding', lambda: check(_lib.load().vdm_timestep_embedding(
        ptr(t_frame), t_frame.numel(), dim, float(max_period), ptr(out), stream()), 'vdm_timestep_embedding'))


def rpe_hidden(e_t, frame_indices, wd, bd, B, T, Cc, out, et_offsets=None, n_blocks=1):
    """First layer of the RPE nets for n_blocks attention blocks at once (wd, bd, out hold 3*n_blocks nets)."""
    _timed('rpe_hidden', lambda: check(_lib.load().vdm_rpe_hidden(
        e_t.data_ptr(), e_t.stride(0), ptr(et_offsets), n_blocks, ptr(frame_indices), ptr(wd), ptr(bd), B, T, Cc,
        ptr(out), dt(out.dtype), stream()), 'vdm_rpe_hidden'), nbytes=_nbytes(out))


def attn_temporal(qkv, r_q, r_k, r_v, mask, pad_interact, B, T, HW, heads, hd, out):
    _timed('attn_temporal', lambda: check(_lib.load().vdm_attn_temporal(
        ptr(qkv), ptr(r_q), ptr(r_k), ptr(r_v), ptr(mask), int(pad_interact), B, T, HW, heads, hd, ptr(out),
        dt(out.dtype), stream()), 'vdm_attn_temporal'), flops=2.0 * 5 * B * HW * heads * T * T * hd,
           nbytes=_nbytes(qkv, out))


def rpe_lookup(tables, frame_indices, B, T, Cc, alpha, beta, gamma, out):
    """Lookup-table RPE (use_rpe_net=False): out [3][B*T*T][C] from tables [3][2*beta+1][C]."""
    _timed('rpe_lookup', lambda: check(_lib.load().vdm_rpe_lookup(
        ptr(tables), ptr(frame_indices), B, T, Cc, tables.shape[1], float(alpha), float(beta), float(gamma), ptr(out),
        stream()), 'vdm_rpe_lookup'), nbytes=_nbytes(out))


def rpe_expand(r_q, r_k, r_v, B, T, heads, hd, gpt, bq, bk, bv, bias=None, n_blocks=1, r_block_stride=0,
               zero_fill=True, qk_block_stride=0):
    _timed('rpe_expand', lambda: check(_lib.load().vdm_rpe_expand(
        ptr(r_q), ptr(r_k), ptr(r_v), ptr(bias), n_blocks, r_block_stride, qk_block_stride, B, T, heads, hd, gpt, int(zero_fill), ptr(bq), ptr(bk),
        ptr(bv), stream()), 'vdm_rpe_expand'), nbytes=_nbytes(bq, bk, bv))


def attn_temporal_tc(qkv, sk, sq, mask, pad_interact, B, T, HW, heads, hd, gpt, pm, pv):
    _timed('attn_temporal_tc', lambda: check(_lib.load().vdm_attn_temporal_tc(
        ptr(qkv), ptr(sk), ptr(sq), ptr(mask), int(pad_interact), B, T, HW, heads, hd, gpt, ptr(pm), ptr(pv),
        stream()), 'vdm_attn_temporal_tc'), flops=4.0 * B * HW * heads * T * T * hd,
           nbytes=_nbytes(qkv, sk, sq, pm, pv))


def rpe_pack(r_q, r_k, r_v, B, T, heads, hd, rq, rk, rv, bias=None, n_blocks=1, r_block_stride=0,
             qk_block_stride=0, v_block_stride=0):
    """fp32 R tables -> the fused temporal kernel's fragment-major bf16 operands (B*T*heads*hd*32 elements per table)."""
    _timed('rpe_pack', lambda: check(_lib.load().vdm_rpe_pack(
        ptr(r_q), ptr(r_k), ptr(r_v), ptr(bias), n_blocks, r_block_stride, B, T, heads, hd, ptr(rq), ptr(rk),
        ptr(rv), qk_block_stride, v_block_stride, stream()), 'vdm_rpe_pack'), nbytes=_nbytes(rq, rk, rv))


def attn_temporal_fused_smem(T, hd, t_pad, pixels_per_cta):
    """Shared memory (bytes) of the fused temporal kernel for this shape; -1 = not instantiated."""
    return int(_lib.load().vdm_attn_temporal_fused_smem(T, hd, t_pad, pixels_per_cta))


def attn_temporal_fused(qkv, rq, rk, rv, mask, pad_interact, B, T, HW, heads, hd, t_pad, out, pixels_per_cta=0):
    _timed('attn_temporal_fused', lambda: check(_lib.load().vdm_attn_temporal_fused(
        ptr(qkv), ptr(rq), ptr(rk), ptr(rv), ptr(mask), int(pad_interact), B, T, HW, heads, hd, t_pad, pixels_per_cta,
        ptr(out), stream()), 'vdm_attn_temporal_fused'), flops=10.0 * B * HW * heads * T * T * hd,
           nbytes=_nbytes(qkv, out))


def attn_spatial(qkv, n_img, L, heads, hd, out):
    _timed('attn_spatial', lambda: check(_lib.load().vdm_attn_spatial(
        ptr(qkv), dt(qkv.dtype), n_img, L, heads, hd, ptr(out), dt(out.dtype), stream()), 'vdm_attn_spatial'),
           flops=4.0 * n_img * heads * L * L * hd, nbytes=_nbytes(qkv, out))


def attn_weights_mean(qkv, n_outer, n_inner, outer_stride, inner_stride, seq_stride, L, heads, hd, out, r_q=None,
                      r_k=None, mask=None, pad_interact=True):
    """Head-averaged attention maps for return_attn_weights=True (logging only)."""
    _timed('attn_weights_mean', lambda: check(_lib.load().vdm_attn_weights_mean(
        ptr(qkv), dt(qkv.dtype), n_outer, n_inner, outer_stride, inner_stride, seq_stride, L, heads, hd, ptr(r_q),
        ptr(r_k), ptr(mask), int(pad_interact), ptr(out), stream()), 'vdm_attn_weights_mean'))


_F32, _I64 = torch.float32, torch.int64


def check_timesteps():
    """Raise the reference's IndexError if a completed sampler-family launch saw a timestep outside its schedule
    tables (the kernels clamp and record instead of reading out of bounds).  Asynchronous like any CUDA error:
    called at the start of every sampler-family op and after explicit synchronisation points."""
    if _lib.load().vdm_sampler_error():
        raise IndexError('timestep outside [0, num_timesteps) reached a libvdm sampler kernel (pass SpacedDiffusion '
                         'indices, not original timesteps)')


def sampler_step(mode, x, eps, noise, t, tables, clip_denoised=True, eta=0.0, sample=None, pred_xstart=None,
                 mean=None):
    check_timesteps()
    B = x.shape[0]
    per_batch = x.numel() // B
    if sample is None:
        sample = torch.empty_like(x)
    _timed('sampler_step', lambda: check(_lib.load().vdm_sampler_step(
        mode, ptr(x, _F32), ptr(eps, _F32), ptr(noise, _F32), ptr(t, _I64), ptr(tables, _F32), tables.shape[1], B,
        per_batch, int(clip_denoised), float(eta), ptr(sample, _F32), ptr(pred_xstart, _F32), ptr(mean, _F32), stream()),
        'vdm_sampler_step'), nbytes=_nbytes(x, eps, noise, sample, pred_xstart, mean))
    return sample


def q_sample(x0, noise, t, tables, out=None):
    check_timesteps()
    B = x0.shape[0]
    if out is None:
        out = torch.empty_like(x0)
    check(_lib.load().vdm_q_sample(ptr(x0, _F32), ptr(noise, _F32), ptr(t, _I64), ptr(tables, _F32), tables.shape[1], B,
                                   x0.numel() // B, ptr(out, _F32), stream()), 'vdm_q_sample')
    return out


def lincomb(op, a, b, t, tables, row_a, row_b=0, out=None):
    """Per-row linear combination with schedule coefficients gathered on the device (include/vdm.h: vdm_lincomb)."""
    check_timesteps()
    B = a.shape[0]
    if out is None:
        out = torch.empty_like(a)
    check(_lib.load().vdm_lincomb(op, ptr(a, _F32), ptr(b, _F32), ptr(t, _I64), ptr(tables, _F32), tables.shape[1],
                                  row_a, row_b, B, a.numel() // B, ptr(out, _F32), stream()), 'vdm_lincomb')
    return out


def vb_terms(x0, x_t, eps, noise, t, tables, latent_mask, clip_denoised, acc):
    check_timesteps()
    B, F = x0.shape[0], x0.shape[1]
    per_frame = x0.numel() // (B * F)
    check(_lib.load().vdm_vb_terms(ptr(x0, _F32), ptr(x_t, _F32), ptr(eps, _F32), ptr(noise, _F32), ptr(t, _I64),
                                   ptr(tables, _F32), tables.shape[1], ptr(latent_mask, _F32), B, F, per_frame,
                                   int(clip_denoised), ptr(acc, torch.float64), stream()), 'vdm_vb_terms')


def prior_bpd(x0, tables, latent_mask, acc):
    B, F = x0.shape[0], x0.shape[1]
    check(_lib.load().vdm_prior_bpd(ptr(x0, _F32), ptr(tables, _F32), tables.shape[1], ptr(latent_mask, _F32), B, F,
                                    x0.numel() // (B * F), ptr(acc, torch.float64), stream()), 'vdm_prior_bpd')
