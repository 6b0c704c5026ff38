"""ctypes binding of libvdm.so (the C ABI declared in include/vdm.h).

There is deliberately no fallback: if the CUDA extension is missing or a call fails,
everything here raises.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('VDM_LIB') or os.path.join(_HERE, 'libvdm.so')    # VDM_LIB: diagnostics builds (make TRACE=1)

F32, BF16, F64, I64, F16 = 0, 1, 2, 3, 4
TAB = dict(SQRT_RECIP_ACP=0, SQRT_RECIPM1_ACP=1, POST_C1=2, POST_C2=3, MODEL_LOGVAR=4, MODEL_VAR=5, ACP=6,
           ACP_PREV=7, POST_LOGVAR=8, SQRT_ACP=9, SQRT_1M_ACP=10, LOG_1M_ACP=11, POST_VAR=12, RECIP_POST_C1=13,
           POST_C2_DIV_C1=14, ACP_NEXT=15)
TAB_COUNT = 16

EXPORTS = ['vdm_version', 'vdm_last_error_string', 'vdm_launch_count', 'vdm_gemm_set_trace', 'vdm_gemm', 'vdm_gemm_fused_norm_supported', 'vdm_gemm_img_done_supported', 'vdm_gn_stats', 'vdm_gn_stats_t', 'vdm_gn_apply', 'vdm_gn_coef',
           'vdm_gn_temporal', 'vdm_gn_temporal_t', 'vdm_add_spatial_encoding', 'vdm_add_spatial_encoding_t', 'vdm_cond_mix', 'vdm_stage_inputs', 'vdm_map_timesteps', 'vdm_timestep_embedding', 'vdm_rpe_hidden',
           'vdm_attn_temporal', 'vdm_rpe_lookup', 'vdm_rpe_expand', 'vdm_attn_temporal_tc', 'vdm_rpe_pack', 'vdm_attn_temporal_fused', 'vdm_attn_temporal_fused_smem', 'vdm_attn_temporal_fused_set_trace', 'vdm_attn_spatial', 'vdm_attn_weights_mean', 'vdm_sampler_error', 'vdm_sampler_step', 'vdm_q_sample', 'vdm_lincomb',
           'vdm_vb_terms', 'vdm_prior_bpd']

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [('dtype', _i32), ('taps', _i32), ('a1_mode', _i32), ('n_img', _i32), ('H', _i32), ('W', _i32),
                ('C1', _i32), ('C2', _i32), ('N', _i32), ('a1', _vp), ('a2', _vp), ('w', _vp), ('bias', _vp),
                ('rowbias', _vp), ('ld_rowbias', _i32), ('residual', _vp), ('ld_res', _i32), ('out_f32', _vp),
                ('out_bf16', _vp), ('ld_out', _i32), ('ld_out_bf16', _i32), ('out_nchw', _i32),
                ('out_silu_f32', _vp), ('lda1', _i32), ('w_group_tiles', _i32), ('stats_out', _vp), ('n_prob', _i32),
                ('prob_a_cols', _i32), ('prob_w_rows', _i64), ('prob_out_stride', _i64), ('a1_coef', _vp),
                ('a1_act', _i32), ('a2b', _vp), ('C2b', _i32), ('a2_dtype', _i32), ('io_dtype', _i32), ('img_done', _vp), ('a1_raw_dtype', _i32)]


class GnApplyArgs(C.Structure):
    _fields_ = [('src1', _vp), ('C1', _i32), ('src2', _vp), ('C2', _i32), ('n_img', _i32), ('H', _i32), ('W', _i32),
                ('stats1', _vp), ('stats2', _vp), ('stats_dtype', _i32), ('stats2_dtype', _i32), ('gamma', _vp), ('beta', _vp), ('scale_shift', _vp), ('ld_ss', _i32), ('silu', _i32),
                ('out_mode', _i32), ('out_dtype', _i32), ('out', _vp), ('src1_dtype', _i32), ('out_raw', _vp), ('out_f32_copy', _vp), ('copy_dtype', _i32), ('wait_done', _vp), ('wait_count', C.c_uint32)]


_lib = None


def load():
    """Load libvdm.so (built by `__graft_entry__.build()` / `make -C video_diffusion_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"`. '
                           'There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    lib.vdm_version.restype = _i32
    lib.vdm_last_error_string.restype = C.c_char_p
    lib.vdm_launch_count.restype = _i64
    lib.vdm_sampler_error.restype = _i32
    lib.vdm_sampler_error.argtypes = []
    sig = {
        'vdm_gemm': [C.POINTER(GemmArgs), _vp],
        'vdm_gn_stats': [_vp, _i32, _i32, _i32, _vp, _vp],
        'vdm_gn_stats_t': [_vp, _i32, _i32, _i32, _i32, _vp, _vp],
        'vdm_gn_temporal_t': [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp],
        'vdm_add_spatial_encoding_t': [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
        'vdm_gn_apply': [C.POINTER(GnApplyArgs), _vp],
        'vdm_gemm_fused_norm_supported': [C.POINTER(GemmArgs)],
        'vdm_gemm_img_done_supported': [C.POINTER(GemmArgs)],
        'vdm_gn_coef': [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp],
        'vdm_gn_temporal': [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp],
        'vdm_add_spatial_encoding': [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
        'vdm_cond_mix': [_vp] * 6 + [_i32] * 5 + [_vp, _i32, _vp, _vp, _vp],
        'vdm_timestep_embedding': [_vp, _i32, _i32, C.c_double, _vp, _vp],
        'vdm_rpe_hidden': [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp],
        'vdm_attn_temporal': [_vp] * 5 + [_i32] * 6 + [_vp, _i32, _vp],
        'vdm_attn_spatial': [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
        'vdm_attn_weights_mean': [_vp, _i32, _i64, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp],
        'vdm_rpe_lookup': [_vp, _vp, _i32, _i32, _i32, _i32, C.c_double, C.c_double, C.c_double, _vp, _vp],
        'vdm_rpe_expand': [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
        'vdm_attn_temporal_tc': [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp],
        'vdm_rpe_pack': [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _vp],
        'vdm_attn_temporal_fused_smem': [_i32, _i32, _i32, _i32],
        'vdm_attn_temporal_fused_set_trace': [_vp],
        'vdm_attn_temporal_fused': [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp],
        'vdm_map_timesteps': [_vp, _vp, _i32, _f32, _vp, _i32, _vp],
        'vdm_stage_inputs': [_vp] * 7 + [_i32, _i32, _i64] + [_vp] * 8,
        'vdm_sampler_step': [_i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _f32, _vp, _vp, _vp, _vp],
        'vdm_q_sample': [_vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp],
        'vdm_lincomb': [_i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp],
        'vdm_vb_terms': [_vp] * 6 + [_i32, _vp, _i32, _i32, _i64, _i32, _vp, _vp],
        'vdm_prior_bpd': [_vp, _vp, _i32, _vp, _i32, _i32, _i64, _vp, _vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _i64 if name == 'vdm_attn_temporal_fused_smem' else _i32
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().vdm_last_error_string().decode(errors='replace')
        raise RuntimeError(f'libvdm {what} failed (rc={rc}): {msg}')


def ptr(t, dtype=None):
    """Device pointer of a tensor (None -> NULL).  The kernels reinterpret raw memory, so the tensor must be a
    contiguous CUDA tensor ON THE CURRENT DEVICE (launches go to that device's current stream) and, when the entry
    point fixes the element type, of exactly that dtype -- anything else raises instead of computing garbage."""
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        raise ValueError('libvdm needs contiguous CUDA tensors')
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f'libvdm: tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()} '
                           '(wrap the call in torch.cuda.device(tensor.device))')
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f'libvdm: expected a {dtype} tensor, got {t.dtype}')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(load().vdm_launch_count())


def dt(code_or_dtype):
    if isinstance(code_or_dtype, int):
        return {BF16: torch.bfloat16, F16: torch.float16}.get(code_or_dtype, torch.float32)
    return {torch.bfloat16: BF16, torch.float16: F16}.get(code_or_dtype, F32)
