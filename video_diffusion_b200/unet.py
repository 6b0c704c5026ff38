"""B200-native video U-Net behind the reference's model API.

`UNetModel` / `UNetVideoModel` / `CondMargVideoModel` take the reference's constructor kwargs
(improved_diffusion/unet.py:564-586, 880-896, 932-947), expose parameters under the reference's
`state_dict` key names (so `load_state_dict(reference_checkpoint['state_dict'])` works unchanged)
and keep its forward signatures and return values (unet.py:898-912, 949-1026).  No sub-module has a
forward of its own: the model pre-packs its weights for the kernels (OIHW -> [Cout][tap][Cin] in the
compute dtype, fused skip projections, batched embedding projections) and executes one flat plan
of libvdm launches over channels-last activations, optionally replayed as a CUDA graph.

compute_dtype = torch.bfloat16 (default): tcgen05 tensor-core GEMMs, bf16 operands, fp32 accumulate,
fp32 residual stream / GroupNorm statistics / softmax.  compute_dtype = torch.float32: the
reference-accuracy mode (fp32 SIMT kernels everywhere).

Unsupported configurations raise NotImplementedError -- there is no PyTorch / CPU fallback.
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib, ops


def _uniform(shape, bound):
    return torch.empty(shape).uniform_(-bound, bound)


def fold_upsample_weights(w):
    """(O, I, 3, 3) conv weights -> [4 parities][O][4*I] weights of the equivalent 2x2 convolutions on the
    low-resolution input of `F.interpolate(x, 2, 'nearest')` followed by the 3x3 conv (unet.py:63-72).
    Output pixel (2y+a, 2x+b) reads upsampled rows 2y+a-1..2y+a+1, i.e. low-res rows {y-1, y, y} (a=0) or
    {y, y, y+1} (a=1): the three row taps collapse to two with weights (w0, w1+w2) or (w0+w1, w2); columns
    alike.  Parity index = 2a+b, tap index = 2i+j, K index = tap*I + c."""
    w = w.float()
    rows = [torch.stack([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], dim=2),      # a = 0 : (O, I, 2, 3)
            torch.stack([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], dim=2)]      # a = 1
    out = []
    for a in range(2):
        r = rows[a]
        cols = [torch.stack([r[..., 0], r[..., 1] + r[..., 2]], dim=3),     # b = 0 : (O, I, 2, 2)
                torch.stack([r[..., 0] + r[..., 1], r[..., 2]], dim=3)]     # b = 1
        for b in range(2):
            out.append(cols[b].permute(0, 2, 3, 1).reshape(w.shape[0], -1))  # (O, [i][j][c])
    return torch.stack(out, dim=0).reshape(4 * w.shape[0], -1).contiguous()


class UNetModel(nn.Module):
    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False, use_spatial_encoding=False,
                 image_size=None, temporal_augment_type=None, use_rpe_net=False, bucket_params=None,
                 allow_interactions_between_padding=False, compute_dtype=None):
        super().__init__()
        if dims != 2 or not conv_resample or num_classes is not None:
            raise NotImplementedError('only dims=2, conv_resample=True, num_classes=None are supported')
        if bucket_params is None:   # the reference asserts this too (unet.py:423-427)
            raise AssertionError('rp_alpha / rp_beta / rp_gamma must be set')
        self.use_rpe_net, self.bucket_params = bool(use_rpe_net), dict(bucket_params)
        if num_heads_upsample not in (-1, num_heads):
            raise NotImplementedError('num_heads_upsample != num_heads')
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self._cond_mode = 2          # input-mix mode of vdm_cond_mix: plain 3-channel frames unless a subclass says otherwise
        self.use_frame_encoding = getattr(self, 'use_frame_encoding', False)      # set by UNetVideoModel before this ctor
        self.enforce_position_invariance = getattr(self, 'enforce_position_invariance', False)
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, tuple(attention_resolutions)
        self.dropout, self.channel_mult, self.num_heads = dropout, tuple(channel_mult), num_heads
        self.use_scale_shift_norm, self.image_size = use_scale_shift_norm, image_size
        self.allow_interactions_between_padding = allow_interactions_between_padding
        self.num_classes, self.use_checkpoint, self.conv_resample = None, use_checkpoint, conv_resample
        self.compute_dtype = compute_dtype or torch.bfloat16
        self.use_cuda_graph = True
        self.bf16_intermediate = True     # bf16 mode: keep ResBlock conv1 outputs in bf16 only
        self.overlap_rpe_tables = True    # RPE tables on a side stream, concurrent with the first U-Net blocks
        self.temporal_tensor_cores = True # bf16 mode: RPE terms as grouped GEMMs + mma.sync attention core
        # bf16 mode: the whole temporal attention (RPE score terms, q.k^T, softmax, P.V, attn.R_v) as ONE kernel
        self.fused_temporal = os.environ.get('VDM_FUSED_TEMPORAL', '1') != '0'
        # fp16 stream: the attention qkv projections read the normalised fp16 copy directly (no bf16 operand copy)
        self.qkv_from_stream = os.environ.get('VDM_QKV_F16', '1') != '0'
        self.stage_inputs_fused = os.environ.get('VDM_STAGE_INPUTS', '1') != '0'   # one launch for the per-call input copies
        self._borrow_output = False
        self.fuse_head_norm = os.environ.get('VDM_FUSE_HEAD', '1') != '0'   # out-head GroupNorm + SiLU inside the head conv
        self.temporal_pixels_per_cta = int(os.environ.get('VDM_TEMPORAL_PT', '0'))   # 0 = heuristic
        # bf16 mode: GroupNorm-apply + SiLU inside the conv's operand path (transform warps of the halo kernels).  Correct
        # and bit-exact against the standalone pass, but measured SLOWER on B200 (every activation element
        # is re-transformed 4.5 times, which binds on the special-function units and shared-memory bandwidth): off.
        self.fuse_norm = False
        # bf16 mode: the residual stream (every ResBlock / attention / resampling output) is stored in fp16 -- 11-bit
        # mantissa, 4x finer than the bf16 rounding of the GEMM operands, saturating conversions -- instead of fp32:
        # half the HBM bytes of every stream write, residual read and GroupNorm-apply read.  Accumulation, residual adds
        # and GroupNorm statistics stay fp32.  torch.float32 restores the fp32 stream.
        self.stream_dtype = torch.float16
        # bf16 mode: the U-Net body runs on this many groups of videos in parallel streams (see _body_split)
        self.micro_batches = int(os.environ.get('VDM_MICRO_BATCHES', '2'))
        # opt-in (VDM_MB_JOIN_HW=256): micro-batches join at or below this many pixels per image -- the 16x16 / 8x8
        # levels then run on the whole batch (their GEMMs are short: half a batch leaves the last wave of tiles nearly
        # empty), the 64x64 / 32x32 levels per group.  Measured slower than splitting everywhere (12.75-12.83 vs
        # 12.55-12.67 ms, profiles/bench_r3d_*): what the second stream hides at the small levels is worth more
        self.micro_batch_join_hw = int(os.environ.get('VDM_MB_JOIN_HW', '0'))
        self.identity_res_min_hw = int(os.environ.get('VDM_IDENTITY_RES_MIN_HW', '4096'))   # x + h as an identity K range from this H*W on
        self.pipeline_norm = os.environ.get('VDM_PIPELINE_NORM', '0') != '0'     # out_layers GroupNorm-apply beside conv1
        # proj_out's residual as an identity K range (like conv2 at 64x64): measured SLOWER (0.84 vs 0.71 ms for the 22
        # launches; these short-K linears are bound by operand delivery, not by the epilogue): off
        self.proj_identity = os.environ.get('VDM_PROJ_IDENTITY', '0') != '0'
        self.time_embed_dim = E = model_channels * 4
        if model_channels % 64 or (model_channels // num_heads) % 4:
            raise NotImplementedError('model_channels must be a multiple of 64')

        # ---- enumerate modules exactly like the reference constructor (unet.py:605-749) ----
        ch = model_channels
        self._lin('time_embed.0', ch, E)
        self._lin('time_embed.2', E, E)
        self.plan = [dict(kind='conv_in', p='input_blocks.0.0', group='input_blocks.0')]
        self._conv('input_blocks.0.0', in_channels, ch, 3)
        chans, ds, idx = [ch], 1, 1
        self.n_blocks_before_attn = None
        first_attn = None
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                if ds in self.attention_resolutions and self.n_blocks_before_attn is None:
                    self.n_blocks_before_attn = idx
                    first_attn = (ds, ch)
                self._res(f'input_blocks.{idx}.0', ch, mult * model_channels, f'input_blocks.{idx}')
                ch = mult * model_channels
                if ds in self.attention_resolutions:
                    self._attn(f'input_blocks.{idx}.1', ch, f'input_blocks.{idx}')
                chans.append(ch)
                idx += 1
            if level != len(channel_mult) - 1:
                self._conv(f'input_blocks.{idx}.0.op', ch, ch, 3)
                self.plan.append(dict(kind='down', p=f'input_blocks.{idx}.0.op', group=f'input_blocks.{idx}', C=ch))
                chans.append(ch)
                ds *= 2
                idx += 1
        if self.n_blocks_before_attn is None:
            self.n_blocks_before_attn = idx
            first_attn = (ds, ch)
        if use_spatial_encoding:
            res = image_size // first_attn[0]
            self.spatial_encoding = nn.Parameter(torch.randn(1, first_attn[1], res, res))
        else:
            self.spatial_encoding = None
        self._res('middle_block.0', ch, ch, 'middle_block')
        self._attn('middle_block.1', ch, 'middle_block')
        self._res('middle_block.2', ch, ch, 'middle_block')
        idx = 0
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                group = f'output_blocks.{idx}'
                self._res(f'{group}.0', ch + chans.pop(), model_channels * mult, group, cat=True)
                ch = model_channels * mult
                j = 1
                if ds in self.attention_resolutions:
                    self._attn(f'{group}.1', ch, group)
                    j = 2
                if level and i == num_res_blocks:
                    self._conv(f'{group}.{j}.conv', ch, ch, 3)
                    self.plan.append(dict(kind='up', p=f'{group}.{j}.conv', group=group, C=ch))
                    ds //= 2
                idx += 1
        self._gn('out.0', ch)
        self._conv('out.2', model_channels, out_channels, 3, zero=True)
        self._packed = None
        self._packed_version = -1
        self._workspaces = {}

    # ---- parameter registration under the reference's names ------------------------------------
    def _reg(self, path, tensor):
        mod = self
        parts = path.split('.')
        for name in parts[:-1]:
            if name not in mod._modules:
                mod.add_module(name, nn.Module())
            mod = mod._modules[name]
        mod.register_parameter(parts[-1], nn.Parameter(tensor))

    def _conv(self, p, cin, cout, k, zero=False):
        fan = cin * k * k
        self._reg(p + '.weight', torch.zeros(cout, cin, k, k) if zero else _uniform((cout, cin, k, k), fan ** -0.5))
        self._reg(p + '.bias', torch.zeros(cout) if zero else _uniform((cout,), fan ** -0.5))

    def _lin(self, p, cin, cout, zero=False):
        self._reg(p + '.weight', torch.zeros(cout, cin) if zero else _uniform((cout, cin), cin ** -0.5))
        self._reg(p + '.bias', torch.zeros(cout) if zero else _uniform((cout,), cin ** -0.5))

    def _gn(self, p, c):
        self._reg(p + '.weight', torch.ones(c))
        self._reg(p + '.bias', torch.zeros(c))

    def _res(self, p, cin, cout, group, cat=False):
        self._gn(p + '.in_layers.0', cin)
        self._conv(p + '.in_layers.2', cin, cout, 3)
        self._lin(p + '.emb_layers.1', self.time_embed_dim, 2 * cout if self.use_scale_shift_norm else cout)
        self._gn(p + '.out_layers.0', cout)
        self._conv(p + '.out_layers.3', cout, cout, 3, zero=True)
        if cin != cout:
            self._conv(p + '.skip_connection', cin, cout, 1)
        self.plan.append(dict(kind='res', p=p, group=group, cin=cin, cout=cout, skip=cin != cout, cat=cat))

    def _attn(self, p, c, group):
        for which in ('spatial_attention', 'temporal_attention'):
            q = f'{p}.{which}'
            self._lin(q + '.qkv', c, 3 * c)
            self._lin(q + '.proj_out', c, c, zero=True)
            self._gn(q + '.norm', c)
            if which == 'temporal_attention' and not self.use_rpe_net:
                for r in ('rpe_q', 'rpe_k', 'rpe_v'):      # bucketed lookup table (unet.py:326-328), zero-initialised
                    self._reg(f'{q}.{r}.lookup_table_weight',
                              torch.zeros(2 * self.bucket_params['beta'] + 1, self.num_heads, c // self.num_heads))
            elif which == 'temporal_attention':
                for r in ('rpe_q', 'rpe_k', 'rpe_v'):
                    self._lin(f'{q}.{r}.rpe_net.embed_distances', 3, c)
                    self._lin(f'{q}.{r}.rpe_net.embed_diffusion_time', self.time_embed_dim, c)
                    self._lin(f'{q}.{r}.rpe_net.out', c, c, zero=True)
        self.plan.append(dict(kind='attn', p=p, group=group, C=c))

    # ---- state handling --------------------------------------------------------------------------
    def _invalidate(self):
        self._packed = None
        self._workspaces = {}

    def _param_version(self):
        """Fingerprint of the parameters: in-place version counters (`p.copy_()`, `p.mul_()`, optimiser / EMA updates
        under no_grad bump them) and storage addresses (`p.data = new`, `vector_to_parameters`).  When it changes after
        the first forward the packed bf16 weights and the graphs built on them are redone.  Writes through a `.data`
        alias (`p.data.mul_()`) bypass autograd's version counter and cannot be seen: call `repack()` after those."""
        return hash(tuple((p._version, p.data_ptr()) for p in self.parameters()))

    def repack(self):
        """Re-pack the kernel-side weights from the current parameters (needed only after writes through `.data`
        aliases; every other update is detected, see _param_version)."""
        self._invalidate()
        return self

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._invalidate()
        return out

    def set_compute_dtype(self, dtype):
        assert dtype in (torch.bfloat16, torch.float32)
        self.compute_dtype = dtype
        self._invalidate()
        return self

    @property
    def inner_dtype(self):
        return torch.float32

    @property
    def _sdt(self):
        """dtype of the residual-stream buffers"""
        return self.stream_dtype if self.compute_dtype == torch.bfloat16 else torch.float32

    # ---- weight pre-packing ----------------------------------------------------------------------
    @torch.no_grad()
    def _pack(self):
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('the B200 model only runs on a CUDA device; there is no CPU fallback (call .to("cuda"))')
        adt = self.compute_dtype
        P = {}

        def conv_w(key):     # OIHW -> [O][tap*I + i]
            w = sd[key]
            return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)

        def put(name, t, dtype=torch.float32):
            P[name] = t.to(device=dev, dtype=dtype).contiguous()

        put('te_w0', sd['time_embed.0.weight']); put('te_b0', sd['time_embed.0.bias'])
        put('te_w2', sd['time_embed.2.weight']); put('te_b2', sd['time_embed.2.bias'])
        w_in = sd['input_blocks.0.0.weight']                  # (ch, Cin, 3, 3) -> im2col columns k = tap*Cin + c
        cin = w_in.shape[1]
        if 9 * cin > 64:
            raise NotImplementedError('input conv: at most 7 input channels')
        pad = torch.zeros(w_in.shape[0], 64)
        pad[:, :9 * cin] = w_in.permute(0, 2, 3, 1).reshape(w_in.shape[0], 9 * cin)
        put('in_w', pad, adt); put('in_b', sd['input_blocks.0.0.bias'])
        emb_w, emb_b, rpe_w, rpe_b = [], [], [], []
        emb_off = rpe_off = 0
        for node in self.plan:
            p = node['p']
            if node['kind'] == 'res':
                put(p + '.gn1_w', sd[p + '.in_layers.0.weight']); put(p + '.gn1_b', sd[p + '.in_layers.0.bias'])
                put(p + '.gn2_w', sd[p + '.out_layers.0.weight']); put(p + '.gn2_b', sd[p + '.out_layers.0.bias'])
                put(p + '.w1', conv_w(p + '.in_layers.2.weight'), adt); put(p + '.b1', sd[p + '.in_layers.2.bias'])
                w2, b2 = conv_w(p + '.out_layers.3.weight'), sd[p + '.out_layers.3.bias']
                if node['skip']:
                    ws = sd[p + '.skip_connection.weight']
                    if ws.shape[-1] != 1:
                        raise NotImplementedError('3x3 skip convolution')
                    w2 = torch.cat([w2, ws.reshape(ws.shape[0], -1)], dim=1)
                    b2 = b2 + sd[p + '.skip_connection.bias']
                put(p + '.w2', w2, adt); put(p + '.b2', b2)
                if adt == torch.bfloat16:
                    # conv2 with its second operand range read straight from the fp16 residual stream: the 3x3 columns
                    # stay bf16, the trailing columns -- the 1x1 skip projection, or an identity that turns the
                    # residual add `x + h` (unet.py:198) into one more K block of the same GEMM -- hold IEEE half bits
                    # (vdm_gemm_args.a2_dtype = VDM_F16).  One 16-bit container; the kernel never interprets it.
                    w3 = conv_w(p + '.out_layers.3.weight')
                    tail = (sd[p + '.skip_connection.weight'].reshape(w3.shape[0], -1) if node['skip']
                            else torch.eye(w3.shape[0], device=w3.device))
                    both = torch.cat([w3.to(torch.bfloat16).view(torch.int16), tail.to(torch.float16).view(torch.int16)],
                                     dim=1)
                    P[p + '.w2s'] = both.to(dev).contiguous().view(torch.bfloat16)
                emb_w.append(sd[p + '.emb_layers.1.weight']); emb_b.append(sd[p + '.emb_layers.1.bias'])
                node['emb_off'] = emb_off
                emb_off += emb_w[-1].shape[0]
            elif node['kind'] == 'attn':
                for which in ('temporal_attention', 'spatial_attention'):
                    q = f'{p}.{which}'
                    put(q + '.gn_w', sd[q + '.norm.weight']); put(q + '.gn_b', sd[q + '.norm.bias'])
                    put(q + '.qkv_w', sd[q + '.qkv.weight'], adt); put(q + '.qkv_b', sd[q + '.qkv.bias'])
                    if adt == torch.bfloat16:     # fp16 stream: the normalised copy xn is itself the qkv operand (VDM_F16)
                        put(q + '.qkv_w16', sd[q + '.qkv.weight'], torch.float16)
                    put(q + '.proj_w', sd[q + '.proj_out.weight'], adt); put(q + '.proj_b', sd[q + '.proj_out.bias'])
                    if adt == torch.bfloat16:     # proj_out + identity columns (fp16): `x + proj(a)` as one GEMM, see w2s
                        pw = sd[q + '.proj_out.weight']
                        both = torch.cat([pw.to(torch.bfloat16).view(torch.int16),
                                          torch.eye(pw.shape[0], device=pw.device).to(torch.float16).view(torch.int16)], dim=1)
                        P[q + '.proj_ws'] = both.to(dev).contiguous().view(torch.bfloat16)
                q = p + '.temporal_attention'
                nets = ('rpe_q', 'rpe_k', 'rpe_v')
                if not self.use_rpe_net:
                    put(q + '.rpe_lut', torch.stack([sd[f'{q}.{r}.lookup_table_weight'].reshape(-1, node['C'])
                                                     for r in nets]))
                    continue
                put(q + '.rpe_wd', torch.stack([sd[f'{q}.{r}.rpe_net.embed_distances.weight'] for r in nets]))
                put(q + '.rpe_bd', torch.stack([sd[f'{q}.{r}.rpe_net.embed_distances.bias'] for r in nets]))
                put(q + '.rpe_out_w', torch.cat([sd[f'{q}.{r}.rpe_net.out.weight'] for r in nets]), adt)
                put(q + '.rpe_out_b', torch.stack([sd[f'{q}.{r}.rpe_net.out.bias'] for r in nets]))
                for r in nets:
                    put(f'{q}.{r}.out_w', sd[f'{q}.{r}.rpe_net.out.weight'], adt)
                    put(f'{q}.{r}.out_b', sd[f'{q}.{r}.rpe_net.out.bias'])
                    rpe_w.append(sd[f'{q}.{r}.rpe_net.embed_diffusion_time.weight'])
                    rpe_b.append(sd[f'{q}.{r}.rpe_net.embed_diffusion_time.bias'])
                node['rpe_off'] = rpe_off
                rpe_off += 3 * node['C']
            elif node['kind'] in ('down', 'up'):
                put(p + '.w', conv_w(p + '.weight'), adt); put(p + '.b', sd[p + '.bias'])
                if node['kind'] == 'up' and adt == torch.bfloat16:
                    put(p + '.wfold', fold_upsample_weights(sd[p + '.weight']), adt)
        put('emb_w', torch.cat(emb_w)); put('emb_b', torch.cat(emb_b))
        if rpe_w:
            put('rpe_t_w', torch.cat(rpe_w)); put('rpe_t_b', torch.cat(rpe_b))
        if adt == torch.bfloat16:
            put('emb_w_a', torch.cat(emb_w), adt)
            if rpe_w:
                put('rpe_t_w_a', torch.cat(rpe_w), adt)
        put('out_gn_w', sd['out.0.weight']); put('out_gn_b', sd['out.0.bias'])
        put('out_w', conv_w('out.2.weight'), adt); put('out_b', sd['out.2.bias'])
        if self.spatial_encoding is not None:
            enc = sd['spatial_encoding'][0]                    # (C, res, res) -> [res*res][C]
            put('enc', enc.permute(1, 2, 0).reshape(-1, enc.shape[0]))
        self._n_gn = 2 * sum(n['kind'] == 'res' for n in self.plan) + sum(n['kind'] == 'attn' for n in self.plan) + 1
        self._packed = P

    # ---- execution -------------------------------------------------------------------------------
    class _Workspace:
        def __init__(self, B, F, H, W, dev, stat_capacity, parent=None, lo=0):
            """parent / lo: this is the workspace of one micro-batch (videos lo .. lo+B of the parent's batch): its
            inputs and output are views of the parent's, everything else (activations, statistics) is its own."""
            self.B, self.F, self.H, self.W, self.dev = B, F, H, W, dev
            self.parent, self.lo = parent, lo
            self.bufs = {}
            self.flags = {}          # per-shape dispatch decisions (e.g. which convs take the fused normalisation)
            self.graph = None
            N = B * F
            f32 = torch.float32
            if parent is None:
                self.x = torch.empty(B, F, 3, H, W, device=dev, dtype=f32)
                self.x0 = torch.empty_like(self.x)
                self.obs = torch.empty(B, F, device=dev, dtype=f32)
                self.lat = torch.empty_like(self.obs)
                self.kinda = torch.empty_like(self.obs)
                self.t = torch.empty(B, device=dev, dtype=f32)
                self.t_override = torch.empty(N, device=dev, dtype=f32)
                self.fi = torch.empty(B, F, device=dev, dtype=torch.long)
                self.fi_float = torch.empty(B, F, device=dev, dtype=f32)
                self.out = None
                self.stream = self.done = None
            else:
                self.fi, self.fi_float = parent.fi[lo:lo + B], parent.fi_float[lo:lo + B]
                self.out = parent.out[lo:lo + B]
                self.stream, self.done, self.done2 = torch.cuda.Stream(device=dev), torch.cuda.Event(), torch.cuda.Event()
            self.children, self.mb_fork = [], None
            self.stat_bufs = {}
            self.pool = torch.zeros(stat_capacity, device=dev, dtype=torch.int64)
            self.pool_used = 0
            self.side = self.ev_fork = self.ev_emb = self.ev_join = None
            self.emb_join = self.rpe_join = None

        def buf(self, name, shape, dtype=torch.float32):
            b = self.bufs.get(name)
            if b is None:
                b = self.bufs[name] = torch.empty(shape, device=self.dev, dtype=dtype)
            return b

        def zeros(self, name, shape, dtype=torch.float32):
            b = self.bufs.get(name)
            if b is None:
                b = self.bufs[name] = torch.zeros(shape, device=self.dev, dtype=dtype)
            return b

        def stats(self, name, n_img, C, dtype):
            """Per-(image, channel) sum / sum-of-squares table [n_img][2][C], carved out of one 8-byte pool that a
            single memset clears at the start of every forward."""
            b = self.stat_bufs.get(name)
            if b is None:
                n = n_img * 2 * C
                if self.pool_used + n > self.pool.numel():
                    raise RuntimeError('internal: GroupNorm statistics pool too small')
                b = self.pool[self.pool_used:self.pool_used + n].view(dtype).view(n_img, 2, C)
                self.pool_used += n
                self.stat_bufs[name] = b
            return b

        def counters(self, name, n_img):
            """Per-image completion counters (int32 [n_img]) of a conv whose GroupNorm-apply runs beside it; they live in
            the statistics pool, so the memset at the start of a forward clears them too."""
            b = self.stat_bufs.get(name)
            if b is None:
                n = (n_img + 1) // 2
                if self.pool_used + n > self.pool.numel():
                    raise RuntimeError('internal: GroupNorm statistics pool too small')
                b = self.pool[self.pool_used:self.pool_used + n].view(torch.int32)[:n_img]
                self.pool_used += n
                self.stat_bufs[name] = b
            return b

        def zero_stats(self):
            if self.pool_used:
                self.pool[:self.pool_used].zero_()

        def tap(self, name):
            """An activation buffer by name, for tests / diagnostics: with micro-batches the rows of the groups
            (consecutive videos) are concatenated back."""
            if name in self.bufs or not self.children:
                return self.bufs[name]
            return torch.cat([c.bufs[name] for c in self.children], dim=0)

    # GroupNorm statistics travel with the activation as (tensor, stats): in bf16 mode the producing GEMM's
    # epilogue accumulates them (deterministic fixed-point atomics); otherwise the standalone kernel does (float64).
    def _fused_stats(self, ws, name, n_img, HW, C):
        if self.compute_dtype == torch.bfloat16 and HW % 32 == 0:
            return ws.stats(name + '.st', n_img, C, torch.int64)
        return None

    def _stats_of(self, ws, name, h, st, n_img, HW):
        if st is None:
            st = ws.stats(name + '.st64', n_img, h.shape[1], torch.float64)
            ops.gn_stats(h, n_img, HW, st)
        return st

    def _res_block(self, ws, node, x1, x2, n_img, H, W, emb_out):
        """x1, x2: (tensor, stats) pairs; x2 is the U-Net skip (concatenated along C) or None."""
        P, adt, p = self._packed, self.compute_dtype, node['p']
        Cin, Cout, HW = node['cin'], node['cout'], H * W
        M = n_img * HW
        src1, src2 = x1[0], (x2[0] if x2 is not None else None)
        st1 = self._stats_of(ws, p + '.in1', src1, x1[1], n_img, HW)
        st2 = self._stats_of(ws, p + '.in2', src2, x2[1], n_img, HW) if x2 is not None else None
        a1 = ws.buf(p + '.a1', (M, Cin), adt)
        # Second operand range of conv2.  With the fp16 residual stream the block input itself is that operand (fp16
        # MMAs against fp16 weight columns): the 1x1 skip projection needs no bf16 copy of x, and on the 64x64 level
        # -- where the epilogue's residual read costs more than one extra K block -- `x + h` becomes an identity
        # projection as well.  Otherwise (fp32 stream, fp32 mode) gn_apply writes the raw cast next to the operand.
        stream_a2 = adt == torch.bfloat16 and self._sdt == torch.float16 and (node['skip'] or HW >= self.identity_res_min_hw)
        araw = ws.buf(p + '.araw', (M, Cin), adt) if (node['skip'] and not stream_a2) else None
        # one pass over the block input: normalised+SiLU operand of conv1 (and the raw cast for the 1x1 skip, if needed)
        ops.gn_apply(src1, src2, n_img, H, W, a1, stats1=st1, stats2=st2, gamma=P[p + '.gn1_w'], beta=P[p + '.gn1_b'],
                     silu=True, out_raw=araw)
        off = node['emb_off']
        ss = self.use_scale_shift_norm
        st_h1 = self._fused_stats(ws, p + '.h1', n_img, HW, Cout)
        rb = None if ss else emb_out[:, off:off + Cout]
        if rb is not None and ws.emb_join is not None:
            torch.cuda.current_stream().wait_event(ws.emb_join)
            ws.emb_join = None
        h1_done = None
        if st_h1 is not None and self.bf16_intermediate:
            # conv1's output is only ever consumed by GroupNorm -> SiLU -> bf16: keep it in bf16 (its
            # statistics come from the fp32 accumulators in the epilogue), halving its HBM traffic
            h1 = ws.buf(p + '.h1b', (M, Cout), torch.bfloat16)
            conv1 = dict(n_img=n_img, H=H, W=W, taps=9, bias=P[p + '.b1'], rowbias=rb, out_bf16=h1, stats_out=st_h1)
            # Image-pipelined out_layers: where conv1 runs on a kernel that publishes per-image completion counters,
            # the GroupNorm-apply below is launched as its programmatic dependent and normalises image n as soon as
            # conv1 has finished it -- beside conv1 on the SMs' spare registers, reading the image from L2.
            pipe = ws.flags.get(p + '.pipe1')
            if pipe is None:
                pipe = ws.flags[p + '.pipe1'] = bool(self.pipeline_norm and ops.PROFILE is None and
                                                     ops.gemm_img_done_supported(a1, P[p + '.w1'], Cout, **conv1))
            if pipe:
                h1_done = ws.counters(p + '.h1.done', n_img)
                conv1['img_done'] = h1_done
                if ws.emb_join is not None:      # the side branch joins ahead of conv1: nothing between it and its dependent
                    torch.cuda.current_stream().wait_event(ws.emb_join)
                    ws.emb_join = None
            ops.gemm(a1, P[p + '.w1'], Cout, **conv1)
        else:
            h1 = ws.buf(p + '.h1', (M, Cout))
            ops.gemm(a1, P[p + '.w1'], Cout, n_img=n_img, H=H, W=W, taps=9, bias=P[p + '.b1'], rowbias=rb,
                     out_f32=h1, stats_out=st_h1)
            st_h1 = self._stats_of(ws, p + '.h1', h1, st_h1, n_img, HW)
        if ws.emb_join is not None:      # first consumer of the embedding projections: join the side branch
            torch.cuda.current_stream().wait_event(ws.emb_join)
            ws.emb_join = None
        out = ws.buf(p + '.out', (M, Cout), self._sdt)
        st_out = self._fused_stats(ws, p + '.out', n_img, HW, Cout)
        conv2 = dict(n_img=n_img, H=H, W=W, taps=9, bias=P[p + '.b2'], out_f32=out, stats_out=st_out)
        w2 = P[p + '.w2s'] if stream_a2 else P[p + '.w2']
        if stream_a2:
            conv2.update(a2=src1, a2b=src2, C1=Cout)
        else:
            conv2.update(dict(a2=araw) if node['skip'] else dict(residual=src1))
        scale_shift = emb_out[:, off:off + 2 * Cout] if ss else None
        fuse = ws.flags.get(p + '.fuse2')
        if fuse is None:
            # out_layers (unet.py:185-198): GroupNorm -> (1 + scale, shift) -> SiLU -> conv.  Where the conv runs on a
            # halo kernel, its transform warps apply the first three to the raw bf16 conv1 output inside shared
            # memory -- the normalised activation never exists in HBM; elsewhere the standalone pass produces it.
            fuse = ws.flags[p + '.fuse2'] = bool(self.fuse_norm and h1.dtype == torch.bfloat16 and st_out is not None and
                                                 self._sdt == torch.float32 and
                                        ops.gemm_fused_norm_supported(h1, P[p + '.w2'], Cout, **conv2))
        if fuse:
            coef = ws.buf(p + '.coef2', (n_img, Cout, 2))
            ops.gn_coef(st_h1, None, n_img, HW, P[p + '.gn2_w'], P[p + '.gn2_b'], coef, scale_shift=scale_shift)
            ops.gemm(h1, w2, Cout, a1_coef=coef, a1_act=True, **conv2)
        else:
            a2 = ws.buf(p + '.a2', (M, Cout), adt)
            ops.gn_apply(h1, None, n_img, H, W, a2, stats1=st_h1, gamma=P[p + '.gn2_w'], beta=P[p + '.gn2_b'],
                         scale_shift=scale_shift, silu=True, wait_done=h1_done)
            ops.gemm(a2, w2, Cout, **conv2)
        return out, st_out

    def _tc_temporal_ok(self, T, C, HW):
        heads = self.num_heads
        return (self.compute_dtype == torch.bfloat16 and self.temporal_tensor_cores and heads * T <= 128 and T <= 32
                and C // heads in (32, 64, 96, 128) and (HW % 128 == 0 or HW == 64))

    def _temporal_pt(self, T, C, HW):
        """Pixels per CTA of the fused temporal kernel (0 = the shape does not fit it): 16 halves the R-table traffic
        from L2 (hd = 96 only: larger heads do not fit 16-pixel tiles of q, k, v in shared memory), else 8."""
        if not (self.compute_dtype == torch.bfloat16 and self.fused_temporal and T <= 32):
            return 0
        hd, TP = C // self.num_heads, 24 if T <= 24 else 32
        want = self.temporal_pixels_per_cta or (16 if hd == 96 else 8)
        for pt in (want, 8):
            if HW % pt == 0 and 0 < ops.attn_temporal_fused_smem(T, hd, TP, pt) <= 227 * 1024:
                return pt
        return 0

    def _fused_temporal_ok(self, T, C, HW):
        return self._temporal_pt(T, C, HW) != 0

    def _rpe_tables(self, ws, rpe_et, B, T, H, W):
        """RPE tables of every temporal-attention block, batched: the blocks that share (C, HW) go through ONE
        rpe_hidden launch, ONE grouped output-layer GEMM and ONE expansion into the per-(b, t) GEMM operands
        (unet.py:283-296, 357-378).  They depend only on the timestep embedding and the frame indices, so they are
        computed ahead of the U-Net body.  Returns {block prefix: (bq, bk, bv)}."""
        P, adt, heads = self._packed, self.compute_dtype, self.num_heads
        rows = B * T * T
        tables, groups = {}, {}
        if rows % 128 or not self.use_rpe_net:
            return tables
        h, w = H, W
        for node in self.plan:
            if node['kind'] == 'down':
                h, w = h // 2, w // 2
            elif node['kind'] == 'up':
                h, w = 2 * h, 2 * w
            elif node['kind'] == 'attn' and (self._tc_temporal_ok(T, node['C'], h * w)
                                             or self._fused_temporal_ok(T, node['C'], h * w)):
                groups.setdefault((node['C'], h * w), []).append(node)
        for (C, HW), nodes in groups.items():
            nb, key = len(nodes), f'rpe_group.{C}.{HW}'
            if key + '.wd' not in P:
                qs = [n['p'] + '.temporal_attention' for n in nodes]
                P[key + '.wd'] = torch.cat([P[q + '.rpe_wd'] for q in qs]).contiguous()
                P[key + '.bd'] = torch.cat([P[q + '.rpe_bd'] for q in qs]).contiguous()
                P[key + '.out_w'] = torch.cat([P[q + '.rpe_out_w'] for q in qs]).contiguous()
                P[key + '.out_b'] = torch.stack([P[q + '.rpe_out_b'] for q in qs]).contiguous()
                P[key + '.et_off'] = torch.tensor([n['rpe_off'] for n in nodes], dtype=torch.int32, device=rpe_et.device)
            hid = ws.buf(key + '.hid', (nb * 3 * rows, C), adt)
            ops.rpe_hidden(rpe_et, ws.fi, P[key + '.wd'], P[key + '.bd'], B, T, C, hid, et_offsets=P[key + '.et_off'],
                           n_blocks=nb)
            Rall = ws.buf(key + '.R', (nb * 3 * rows, C))
            ops.gemm(hid, P[key + '.out_w'], C, n_img=nb * 3 * rows, H=1, W=1, taps=1, out_f32=Rall,
                     w_group_tiles=rows // 128)
            if self._fused_temporal_ok(T, C, HW):
                # fused kernel: fragment-major bf16 tables, B*T*heads*hd*32 elements each (vdm_rpe_pack)
                hd = C // heads
                per = B * T * heads * hd * 32
                rqk = ws.buf(key + '.rqk', (nb, 2, per), adt)
                rvp = ws.buf(key + '.rvp', (nb, per), adt)
                ops.rpe_pack(Rall[:rows], Rall[rows:2 * rows], Rall[2 * rows:3 * rows], B, T, heads, hd, rqk[0, 0],
                             rqk[0, 1], rvp, bias=P[key + '.out_b'], n_blocks=nb, r_block_stride=3 * rows * C,
                             qk_block_stride=2 * per)
                for i, n in enumerate(nodes):
                    tables[n['p']] = ('fused', rqk[i], rvp[i])
                continue
            gpt = 1 if HW >= 128 else 128 // HW
            SW, ntg = 128 * gpt, (B * T + gpt - 1) // gpt
            # the operands are block-diagonal over heads (~85 % structural zeros): zeroed once with the workspace,
            # afterwards only the live entries are rewritten
            # bk and bq of a block sit back to back, so q -> Sk and k -> Sq run as ONE two-problem GEMM launch
            bkq = ws.zeros(key + '.bkq', (nb, 2, ntg * SW, C), adt)
            bv = ws.zeros(key + '.bv', (nb, ntg * C, SW), adt)
            ops.rpe_expand(Rall[:rows], Rall[rows:2 * rows], Rall[2 * rows:3 * rows], B, T, heads, C // heads, gpt,
                           bkq[0, 1], bkq[0, 0], bv, bias=P[key + '.out_b'], n_blocks=nb, r_block_stride=3 * rows * C,
                           zero_fill=False, qk_block_stride=2 * ntg * SW * C)
            for i, n in enumerate(nodes):
                tables[n['p']] = (bkq[i], bv[i])
        return tables

    def _attention(self, ws, node, x, B, T, H, W, rpe_et, amask, tables=None, attn_log=None):
        P, adt, p, C = self._packed, self.compute_dtype, node['p'], node['C']
        h = x[0]
        HW, heads = H * W, self.num_heads
        hd = C // heads
        N = B * T
        M = N * HW
        lin = dict(n_img=M, H=1, W=1, taps=1)
        # ---- temporal attention with RPE (unet.py:246-255, 471-540)
        q = p + '.temporal_attention'
        xn = ws.buf(q + '.xn', (M, C), self._sdt)
        # fp16 stream: xn (the normalised residual) doubles as the A operand of the qkv projection, in IEEE half
        # with half weights -- 11 instead of 8 significand bits, and no bf16 copy `xa` written and read back
        f16_qkv = self.qkv_from_stream and adt == torch.bfloat16 and xn.dtype == torch.float16
        qkv_a, qkv_w = (xn, P[q + '.qkv_w16']) if f16_qkv else (None, P[q + '.qkv_w'])
        xa = None if f16_qkv else ws.buf(q + '.xa', (M, C), adt)
        ops.gn_temporal(h, B, T, HW, C, P[q + '.gn_w'], P[q + '.gn_b'], xn, xa)
        if qkv_a is None:
            qkv_a = xa
        tc_path = self._tc_temporal_ok(T, C, HW)
        rows = B * T * T
        pre = tables.get(p) if tables else None
        if pre is None and not self.use_rpe_net:
            # lookup-table RPE: the three tables are gathered by the bucketed frame distances (unet.py:326-347)
            Rall = ws.buf(q + '.Rlut', (3, rows, C))
            bp = self.bucket_params
            ops.rpe_lookup(P[q + '.rpe_lut'], ws.fi, B, T, C, bp['alpha'], bp['beta'], bp['gamma'], Rall)
            R, r_bias = [Rall[0], Rall[1], Rall[2]], None
        elif pre is None:
            hid = ws.buf(q + '.hid', (3, rows, C), adt)
            off = node['rpe_off']
            ops.rpe_hidden(rpe_et[:, off:off + 3 * C], ws.fi, P[q + '.rpe_wd'], P[q + '.rpe_bd'], B, T, C, hid)
            R, r_bias = [], None
            for i, r in enumerate(('rpe_q', 'rpe_k', 'rpe_v')):
                Rn = ws.buf(f'{q}.{r}.R', (rows, C))
                ops.gemm(hid[i], P[f'{q}.{r}.out_w'], C, n_img=rows, H=1, W=1, taps=1, bias=P[f'{q}.{r}.out_b'],
                         out_f32=Rn)
                R.append(Rn)
        att = ws.buf(q + '.att', (M, C), adt)
        fused = self._fused_temporal_ok(T, C, HW)
        if fused:
            # ONE kernel between the qkv projection and proj_out (csrc/attention_temporal_fused.cu)
            TP = 24 if T <= 24 else 32
            qkv = ws.buf(q + '.qkvb', (M, 3 * C), adt)
            ops.gemm(qkv_a, qkv_w, 3 * C, bias=P[q + '.qkv_b'], out_bf16=qkv, **lin)
            if pre is not None:
                # tables of the FULL batch; a micro-batch starts at its first (b, t) group
                g0 = tables.get('__group0__', 0) * heads * hd * 32
                rq, rk, rvp = pre[1][0][g0:], pre[1][1][g0:], pre[2][g0:]
            else:
                rqk = ws.buf(q + '.rqk', (2, N * heads * hd * 32), adt)
                rvp = ws.buf(q + '.rvp', (N * heads * hd * 32,), adt)
                ops.rpe_pack(R[0], R[1], R[2], B, T, heads, hd, rqk[0], rqk[1], rvp)
                rq, rk = rqk[0], rqk[1]
            ops.attn_temporal_fused(qkv, rq, rk, rvp, amask, self.allow_interactions_between_padding, B, T, HW, heads,
                                    hd, TP, att, pixels_per_cta=self._temporal_pt(T, C, HW))
        elif tc_path:
            # RPE terms as pixel-batched GEMMs with per-(b, t) weight groups, the rest on mma.sync
            gpt = 1 if HW >= 128 else 128 // HW          # (b, t) groups per 128-row tile
            tpg = max(1, HW // 128)                      # 128-row tiles per group
            SW, ntg = 128 * gpt, (B * T + gpt - 1) // gpt
            qkv = ws.buf(q + '.qkvb', (M, 3 * C), adt)
            ops.gemm(qkv_a, qkv_w, 3 * C, bias=P[q + '.qkv_b'], out_bf16=qkv, **lin)
            sksq = ws.buf(q + '.sksq', (2, M, SW))
            sk, sq = sksq[0], sksq[1]
            if pre is not None:
                # tables of the FULL batch, [2][groups * SW][C] and [groups * C][SW]; a micro-batch reads its (b, t)
                # groups through offset views (the two-problem GEMM keeps the full batch's problem stride)
                bkq, bv = pre
                g0 = tables.get('__group0__', 0) // gpt
                ntg_all = bkq.shape[1] // SW
                bv = bv[g0 * C:(g0 + ntg) * C]
                ops.gemm(qkv[:, :C], bkq.view(2 * ntg_all * SW, C)[g0 * SW:], SW, out_f32=sksq, w_group_tiles=tpg, C1=C,
                         n_prob=2, prob_a_cols=C, prob_w_rows=ntg_all * SW, prob_out_stride=M * SW, **lin)
            else:
                bq = ws.buf(q + '.bq', (ntg * SW, C), adt)
                bk = ws.buf(q + '.bk', (ntg * SW, C), adt)
                bv = ws.buf(q + '.bv', (ntg * C, SW), adt)
                ops.rpe_expand(R[0], R[1], R[2], B, T, heads, hd, gpt, bq, bk, bv, bias=r_bias)
                ops.gemm(qkv[:, :C], bk, SW, out_f32=sk, w_group_tiles=tpg, C1=C, **lin)
                ops.gemm(qkv[:, C:2 * C], bq, SW, out_f32=sq, w_group_tiles=tpg, C1=C, **lin)
            pm = ws.zeros(q + '.pm', (M, SW), adt)       # padding columns stay zero forever
            pv = ws.buf(q + '.pv', (M, C))
            ops.attn_temporal_tc(qkv, sk, sq, amask, self.allow_interactions_between_padding, B, T, HW, heads, hd,
                                 gpt, pm, pv)
            ops.gemm(pm, bv, C, residual=pv, out_bf16=att, w_group_tiles=tpg, **lin)
        else:
            qkv = ws.buf(q + '.qkv', (M, 3 * C))
            ops.gemm(qkv_a, qkv_w, 3 * C, bias=P[q + '.qkv_b'], out_f32=qkv, **lin)
            ops.attn_temporal(qkv, R[0], R[1], R[2], amask, self.allow_interactions_between_padding, B, T, HW, heads,
                              hd, att)
        if attn_log is not None:
            # logging only (unet.py:464-468): |mean over heads| of the (B*HW, T, T) attention maps, recomputed from
            # q, k, the RPE tables and the frame mask
            amap = torch.empty(B * HW, T, T, device=h.device)
            ops.attn_weights_mean(qkv, B, HW, T * HW * 3 * C, 3 * C, HW * 3 * C, T, heads, hd, amap, r_q=R[0], r_k=R[1],
                                  mask=amask, pad_interact=self.allow_interactions_between_padding)
            attn_log['temporal'].append(amap)
        h2 = ws.buf(q + '.out', (M, C), self._sdt)
        st = self._fused_stats(ws, q + '.out', N, HW, C)
        # + NORMALISED x.  The rows of this GEMM are (image, pixel), so the epilogue statistics are
        # exactly the per-image GroupNorm sums the spatial attention needs next.
        self._proj_out(att, q, xn, h2, st, N, H, W, C)
        st = self._stats_of(ws, q + '.out', h2, st, N, HW)
        # ---- spatial attention (unet.py:258-266)
        q = p + '.spatial_attention'
        xn = ws.buf(q + '.xn', (M, C), self._sdt)
        qkv_a, qkv_w = (xn, P[q + '.qkv_w16']) if f16_qkv else (None, P[q + '.qkv_w'])
        xa = None if f16_qkv else ws.buf(q + '.xa', (M, C), adt)
        ops.gn_apply(h2, None, N, H, W, xa, stats1=st, gamma=P[q + '.gn_w'], beta=P[q + '.gn_b'], copy=xn)
        if qkv_a is None:
            qkv_a = xa
        qkv = ws.buf(q + '.qkv', (M, 3 * C), adt)       # bf16 mode: tensor-core flash kernel on bf16 q, k, v
        if adt == torch.bfloat16:
            ops.gemm(qkv_a, qkv_w, 3 * C, bias=P[q + '.qkv_b'], out_bf16=qkv, **lin)
        else:
            ops.gemm(qkv_a, qkv_w, 3 * C, bias=P[q + '.qkv_b'], out_f32=qkv, **lin)
        att = ws.buf(q + '.att', (M, C), adt)
        ops.attn_spatial(qkv, N, HW, heads, hd, att)
        if attn_log is not None:
            amap = torch.empty(N, HW, HW, device=h.device)
            ops.attn_weights_mean(qkv, N, 1, HW * 3 * C, 0, 3 * C, HW, heads, hd, amap)
            attn_log['spatial'].append(amap)
        h3 = ws.buf(q + '.out', (M, C), self._sdt)
        st3 = self._fused_stats(ws, q + '.out', N, HW, C)
        self._proj_out(att, q, xn, h3, st3, N, H, W, C)
        return h3, st3

    def _proj_out(self, att, q, xn, out, st, N, H, W, C):
        """x + proj_out(a) (unet.py:537-538; x = the NORMALISED input).  With the fp16 stream the residual is
        a second K range of the GEMM -- fp16 MMAs of xn against identity columns -- instead of an epilogue read:
        this short-K linear is bound by its epilogue, not by the tensor cores."""
        P = self._packed
        if self.proj_identity and xn.dtype == torch.float16:
            ops.gemm(att, P[q + '.proj_ws'], C, n_img=N, H=H, W=W, taps=1, bias=P[q + '.proj_b'], a2=xn, out_f32=out,
                     stats_out=st, C1=C)
        else:
            ops.gemm(att, P[q + '.proj_w'], C, n_img=N, H=H, W=W, taps=1, bias=P[q + '.proj_b'], residual=xn, out_f32=out,
                     stats_out=st)

    def _run(self, ws, T_attn, per_frame_t, attn_log=None):
        """The whole forward as a flat sequence of libvdm launches: a prologue on the full batch (conditioning mix,
        timestep-embedding branch, RPE tables), then the U-Net body -- on the whole batch, or on `micro_batches`
        groups of videos in parallel streams (see _body_split)."""
        P, adt = self._packed, self.compute_dtype
        B, F, H, W = ws.B, ws.F, ws.H, ws.W
        N, ch, E = B * F, self.model_channels, self.time_embed_dim
        ws.zero_stats()
        a_in = ws.buf('a_in', (N * H * W, 64), adt)
        t_frame, amask = ws.buf('t_frame', (N,)), ws.buf('amask', (N,))
        ops.cond_mix(ws.x, ws.x0, ws.obs, ws.lat, ws.kinda, ws.t, B, F, H, W, a_in, t_frame, amask,
                     mode=self._cond_mode)
        if per_frame_t:
            t_frame = ws.t_override
        emb_out = ws.buf('emb_out', (N, P['emb_w'].shape[0]))
        rpe_et = ws.buf('rpe_et', (N, P['rpe_t_w'].shape[0])) if self.use_rpe_net else None
        lin = dict(n_img=N, H=1, W=1, taps=1)

        def embedding_branch():
            """timestep embedding MLP -> the two wide projections (unet.py:605-610, 156-160, 283-296)."""
            temb = ws.buf('temb', (N, ch))
            ops.timestep_embedding(t_frame, ch, temb)
            l0, l0s = ws.buf('te_l0', (N, E)), ws.buf('te_l0s', (N, E))
            ops.gemm(temb, P['te_w0'], E, bias=P['te_b0'], out_f32=l0, out_silu=l0s, **lin)
            emb, embs = ws.buf('emb', (N, E)), ws.buf('embs', (N, E))
            ops.gemm(l0s, P['te_w2'], E, bias=P['te_b2'], out_f32=emb, out_silu=embs, **lin)
            if adt == torch.bfloat16:
                # the two wide projections of the embedding (all ResBlock scale/shift vectors, all RPE-net time
                # terms) go through the tensor-core GEMM on bf16 copies of silu(emb) / emb
                embs_b, emb_b = ws.buf('embs_b', (N, E), adt), ws.buf('emb_b', (N, E), adt)
                ops.gn_apply(embs, None, N, 1, 1, embs_b)
                ops.gn_apply(emb, None, N, 1, 1, emb_b)
                ops.gemm(embs_b, P['emb_w_a'], P['emb_w'].shape[0], bias=P['emb_b'], out_f32=emb_out, **lin)
                if self.use_rpe_net:
                    ops.gemm(emb_b, P['rpe_t_w_a'], P['rpe_t_w'].shape[0], bias=P['rpe_t_b'], out_f32=rpe_et, **lin)
            else:
                ops.gemm(embs, P['emb_w'], P['emb_w'].shape[0], bias=P['emb_b'], out_f32=emb_out, **lin)
                if self.use_rpe_net:
                    ops.gemm(emb, P['rpe_t_w'], P['rpe_t_w'].shape[0], bias=P['rpe_t_b'], out_f32=rpe_et, **lin)

        # The embedding path and the RPE tables depend only on the timesteps and frame indices: they run on a side
        # stream (a parallel branch of the CUDA graph) under the input conv and the first block, where their small
        # latency-bound launches would otherwise leave the GPU idle.  Two joins: the embedding projections before
        # the first scale/shift GroupNorm, the tables before the first attention block.
        tables, emb_join, rpe_join = None, None, None
        if self.overlap_rpe_tables and (ops.PROFILE is None or ops.PROFILE_STREAMS is not None):
            main = torch.cuda.current_stream()
            if ws.side is None:
                # high priority: the branch is a chain of small launches that the U-Net body waits for (first scale/shift
                # GroupNorm, first attention block); without it they queue behind the body's persistent GEMM grids
                # (VDM_SIDE_PRIORITY=0: default priority, for A/B runs)
                prio = 0 if os.environ.get('VDM_SIDE_PRIORITY') == '0' else -1
                ws.side = torch.cuda.Stream(device=ws.dev, priority=prio)
                ws.ev_fork, ws.ev_emb, ws.ev_join = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            ws.ev_fork.record(main)
            ws.side.wait_event(ws.ev_fork)
            with torch.cuda.stream(ws.side):
                embedding_branch()
                ws.ev_emb.record(ws.side)
                if T_attn == F and attn_log is None:
                    tables = self._rpe_tables(ws, rpe_et, B, F, H, W)
                ws.ev_join.record(ws.side)
            emb_join, rpe_join = ws.ev_emb, ws.ev_join
        else:
            embedding_branch()
            if T_attn == F and attn_log is None:        # the logging kernel reads the per-block fp32 tables
                tables = self._rpe_tables(ws, rpe_et, B, F, H, W)

        pro = dict(a_in=a_in, amask=amask, emb_out=emb_out, rpe_et=rpe_et, tables=tables or {})
        if ws.out is None:
            ws.out = torch.empty(B, F, self.out_channels, H, W, device=ws.dev)
        n_mb = self._micro_batches_for(ws, attn_log)
        if n_mb == 1:
            ws.emb_join, ws.rpe_join = emb_join, rpe_join
            self._body(ws, pro, T_attn, attn_log)
        else:
            self._body_split(ws, pro, T_attn, n_mb, emb_join, rpe_join)
        return ws.out

    # ---- micro-batches -----------------------------------------------------------------------------
    def _micro_batches_for(self, ws, attn_log):
        n = self.micro_batches
        if (n <= 1 or self.compute_dtype != torch.bfloat16 or attn_log is not None
                or (ops.PROFILE is not None and ops.PROFILE_STREAMS is None)
                or ws.B % n or (ws.B // n * ws.F) % 2):     # (b, t) groups of the RPE tables pair up images at 8x8
            return 1
        return n

    def _body_split(self, ws, pro, T_attn, n_mb, emb_join, rpe_join):
        """The videos of a batch are independent (attention couples the frames of ONE video only), so the U-Net body
        runs on `n_mb` groups of videos in parallel streams -- parallel branches of the CUDA graph.  The persistent
        tensor-core GEMMs of the groups serialise on the SMs' shared memory, which staggers the groups by themselves:
        while one group's GEMM holds the tensor cores, the other group's HBM-bound kernels (GroupNorm-apply, temporal
        GroupNorm, casts) run beside it on the same SMs, instead of leaving the tensor cores idle in between."""
        B, F, H, W = ws.B, ws.F, ws.H, ws.W
        Bm = B // n_mb
        if not ws.children:
            for k in range(n_mb):
                ws.children.append(self._Workspace(Bm, F, H, W, ws.dev, ws.pool.numel() // n_mb + 64, parent=ws,
                                                   lo=k * Bm))
            ws.mb_fork = torch.cuda.Event()
            ws.mb_fork2 = torch.cuda.Event()
        main = torch.cuda.current_stream()
        ws.mb_fork.record(main)
        rows_img = H * W
        first, last = self._deep_range(H, W)
        subs = []
        for k, child in enumerate(ws.children):
            lo, hi = k * Bm * F, (k + 1) * Bm * F
            subs.append(dict(a_in=pro['a_in'][lo * rows_img:hi * rows_img], amask=pro['amask'][lo:hi],
                             emb_out=pro['emb_out'][lo:hi],
                             rpe_et=None if pro['rpe_et'] is None else pro['rpe_et'][lo:hi],
                             tables=pro['tables'], table_group0=lo))
        if first is None:
            for child, sub in zip(ws.children, subs):
                child.stream.wait_event(ws.mb_fork)
                with torch.cuda.stream(child.stream):
                    child.zero_stats()
                    child.emb_join, child.rpe_join = emb_join, rpe_join
                    self._body(child, sub, T_attn, None)
                    child.done.record(child.stream)
            for child in ws.children:
                main.wait_event(child.done)
            return
        # phase A, per group: input conv .. the downsample that enters the joined levels (its output and statistics are
        # row ranges of the parent's buffers)
        states = []
        for child, sub in zip(ws.children, subs):
            child.stream.wait_event(ws.mb_fork)
            with torch.cuda.stream(child.stream):
                child.zero_stats()
                child.emb_join, child.rpe_join = emb_join, rpe_join
                states.append(self._body(child, sub, T_attn, None, stop=first + 1, share=first))
                child.done.record(child.stream)
        for child in ws.children:
            main.wait_event(child.done)
        # phase B, whole batch on the main stream: the joined levels
        st0 = states[0]
        full = ws.bufs[st0['shared_name']]
        full_st = ws.stat_bufs.get(st0['shared_name'] + '.st')
        state = dict(st0, hs=[], x=(full, full_st))
        ws.emb_join, ws.rpe_join = emb_join, rpe_join
        state = self._body(ws, pro, T_attn, None, start=first + 1, stop=last, state=state)
        assert not state['hs']
        ws.mb_fork2.record(main)
        # phase C, per group again: from the upsample that leaves the joined levels to the output head
        xb, hw = state['x'], state['H'] * state['W']
        for k, (child, sub) in enumerate(zip(ws.children, subs)):
            lo, hi = k * Bm * F, (k + 1) * Bm * F
            st = dict(states[k], x=(xb[0][lo * hw:hi * hw], None if xb[1] is None else xb[1][lo:hi]),
                      cur_group=state['cur_group'], in_groups=state['in_groups'],
                      n_groups_done=state['n_groups_done'], H=state['H'], W=state['W'])
            child.stream.wait_event(ws.mb_fork2)
            with torch.cuda.stream(child.stream):
                self._body(child, sub, T_attn, None, start=last, state=st)
                child.done2.record(child.stream)
        for child in ws.children:
            main.wait_event(child.done2)

    def _deep_range(self, H, W):
        """(index of the downsample whose output has <= micro_batch_join_hw pixels, index of the upsample that leaves
        those levels), or (None, None) when the micro-batches never join."""
        first = last = None
        if self.micro_batch_join_hw <= 0:
            return first, last
        h, w = H, W
        for i, node in enumerate(self.plan):
            if node['kind'] == 'down':
                h, w = h // 2, w // 2
                if first is None and h * w <= self.micro_batch_join_hw:
                    first = i
            elif node['kind'] == 'up':
                if first is not None and h * w <= self.micro_batch_join_hw < 4 * h * w:
                    last = i
                h, w = 2 * h, 2 * w
        if first is None or last is None or last <= first + 1:
            return None, None
        return first, last

    def _body(self, ws, pro, T_attn, attn_log, start=0, stop=None, state=None, share=None):
        """input conv -> ... -> output head for the videos of `ws` (the whole batch or one micro-batch).  The plan can
        be run in pieces (micro-batches that join for the small levels, see _body_split): nodes [start, stop), the
        traversal state handed on as a dict; `share` = index of a downsample whose output lives in the PARENT's buffers
        (this group's row range of them)."""
        P, adt = self._packed, self.compute_dtype
        B, F, H, W = ws.B, ws.F, ws.H, ws.W
        N, ch = B * F, self.model_channels
        a_in, amask, emb_out, rpe_et = pro['a_in'], pro['amask'], pro['emb_out'], pro['rpe_et']
        tables = dict(pro['tables'])
        tables['__group0__'] = pro.get('table_group0', 0)
        stop = len(self.plan) if stop is None else stop
        shared_name = None

        # activations travel as (tensor, per-channel GroupNorm statistics or None)
        hs, x, cur_group, n_groups_done = [], None, None, 0
        in_groups = True
        if state is not None:
            hs, x, cur_group, n_groups_done = state['hs'], state['x'], state['cur_group'], state['n_groups_done']
            in_groups, H, W = state['in_groups'], state['H'], state['W']

        def close_group():
            nonlocal x, n_groups_done
            if in_groups:
                hs.append(x)
                n_groups_done += 1
                frame_enc = self.use_frame_encoding
                if n_groups_done == self.n_blocks_before_attn and ('enc' in P or frame_enc):
                    hn = ws.buf('h_enc', tuple(x[0].shape), self._sdt)
                    femb = None
                    if frame_enc:     # sinusoid of the (optionally centred) frame indices, period 10 T (unet.py:914-926)
                        femb = ws.buf('frame_emb', (N, x[0].shape[1]))
                        ops.timestep_embedding(ws.fi_float, x[0].shape[1], femb, max_period=self.T * 10)
                    ops.add_spatial_encoding(x[0], P.get('enc'), hn, N, H * W, x[0].shape[1], frame_emb=femb)
                    x = (hn, None)

        for idx in range(start, stop):
            node = self.plan[idx]
            if node['group'] != cur_group:
                if cur_group is not None:
                    close_group()
                cur_group = node['group']
                if cur_group.startswith('middle') or cur_group.startswith('output'):
                    in_groups = False
            kind, p = node['kind'], node['p']
            if kind == 'conv_in':
                h = ws.buf('h_in', (N * H * W, ch), self._sdt)
                st = self._fused_stats(ws, 'h_in', N, H * W, ch)
                ops.gemm(a_in, P['in_w'], ch, n_img=N, H=H, W=W, taps=1, bias=P['in_b'], out_f32=h, stats_out=st)
                x = (h, st)
            elif kind == 'res':
                skip = hs.pop() if node['cat'] else None
                x = self._res_block(ws, node, x, skip, N, H, W, emb_out)
            elif kind == 'attn':
                if T_attn != F:
                    raise NotImplementedError('cross_frame_attention=False')
                if ws.rpe_join is not None:
                    torch.cuda.current_stream().wait_event(ws.rpe_join)
                    ws.rpe_join = None
                x = self._attention(ws, node, x, B, F, H, W, rpe_et, amask, tables, attn_log)
            elif kind == 'down':
                C = node['C']
                hw2 = (H // 2) * (W // 2)
                if idx == share:
                    par = ws.parent
                    shared_name = p + '.out'
                    out = par.buf(shared_name, (par.B * F * hw2, C), self._sdt)[ws.lo * F * hw2:(ws.lo + B) * F * hw2]
                else:
                    out = ws.buf(p + '.out', (N * hw2, C), self._sdt)
                if adt == torch.bfloat16:
                    planes = ws.buf(p + '.planes', (N * H * W, C), adt)
                    ops.gn_apply(x[0], None, N, H, W, planes, out_mode=2)
                    a1 = planes
                else:
                    a1 = x[0]
                if idx == share:
                    st = self._fused_stats(ws.parent, p + '.out', ws.parent.B * F, hw2, C)
                    st = None if st is None else st[ws.lo * F:(ws.lo + B) * F]
                else:
                    st = self._fused_stats(ws, p + '.out', N, hw2, C)
                ops.gemm(a1, P[p + '.w'], C, n_img=N, H=H // 2, W=W // 2, taps=9, a1_mode=1, bias=P[p + '.b'],
                         out_f32=out, C1=C, stats_out=st)
                x, H, W = (out, st), H // 2, W // 2
            elif kind == 'up':
                C = node['C']
                out = ws.buf(p + '.out', (N * 4 * H * W, C), self._sdt)
                fold = adt == torch.bfloat16 and C % 128 == 0
                st = self._fused_stats(ws, p + '.out', N, H * W if fold else 4 * H * W, C)
                if fold:
                    # nearest-x2 + 3x3 conv == four 2x2 parity convs on the low-res input (4/9 of the FLOPs)
                    lo = ws.buf(p + '.lo', (N * H * W, C), adt)
                    ops.gn_apply(x[0], None, N, H, W, lo)
                    ops.gemm(lo, P[p + '.wfold'], C, n_img=N, H=2 * H, W=2 * W, taps=4, a1_mode=3, bias=P[p + '.b'],
                             out_f32=out, stats_out=st, C1=C)
                elif adt == torch.bfloat16:
                    up = ws.buf(p + '.up', (N * 4 * H * W, C), adt)
                    ops.gn_apply(x[0], None, N, H, W, up, out_mode=1)
                    ops.gemm(up, P[p + '.w'], C, n_img=N, H=2 * H, W=2 * W, taps=9, bias=P[p + '.b'], out_f32=out,
                             stats_out=st)
                else:
                    ops.gemm(x[0], P[p + '.w'], C, n_img=N, H=2 * H, W=2 * W, taps=9, a1_mode=2, bias=P[p + '.b'],
                             out_f32=out, C1=C)
                x, H, W = (out, st), 2 * H, 2 * W
        if stop < len(self.plan):
            return dict(hs=hs, x=x, cur_group=cur_group, n_groups_done=n_groups_done, in_groups=in_groups, H=H, W=W,
                        shared_name=shared_name)
        for ev in (ws.rpe_join, ws.emb_join):       # a model without attention / scale-shift: still join the side branch
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
        ws.rpe_join = ws.emb_join = None
        h = x[0]
        st = self._stats_of(ws, 'out', h, x[1], N, H * W)
        if (self.fuse_head_norm and adt == torch.bfloat16 and h.dtype == torch.float16 and self.out_channels <= 8
                and ch % 32 == 0 and ch <= 256 and W % 16 == 0 and H % 8 == 0):
            # GroupNorm-apply + SiLU while the head conv stages its tile (csrc/conv_small_n.cu): no bf16 copy of the
            # top-level stream is written
            coef = ws.buf('out.coef', (N, ch, 2))
            ops.gn_coef(st, None, N, H * W, P['out_gn_w'], P['out_gn_b'], coef)
            ops.gemm(h, P['out_w'], self.out_channels, n_img=N, H=H, W=W, taps=9, bias=P['out_b'], out_f32=ws.out,
                     out_nchw=True, a1_coef=coef, a1_act=True)
            return
        a = ws.buf('out.a', (N * H * W, ch), adt)
        ops.gn_apply(h, None, N, H, W, a, stats1=st, gamma=P['out_gn_w'], beta=P['out_gn_b'], silu=True)
        ops.gemm(a, P['out_w'], self.out_channels, n_img=N, H=H, W=W, taps=9, bias=P['out_b'], out_f32=ws.out,
                 out_nchw=True)

    def _execute(self, x, *args, **kwargs):
        """Every libvdm launch goes to the current stream of the current device: make that the tensors' device."""
        if not x.is_cuda:
            raise RuntimeError('the B200 model needs CUDA tensors; there is no CPU fallback')
        with torch.cuda.device(x.device):
            return self._execute_on_device(x, *args, **kwargs)

    def _execute_on_device(self, x, x0, obs, lat, kinda, t, frame_indices, T_attn, per_frame_t=None, clone=True,
                           attn_log=None):
        if self.training:
            raise NotImplementedError('the B200 model is inference-only: call .eval()')
        if not x.is_cuda:
            raise RuntimeError('the B200 model needs CUDA tensors; there is no CPU fallback')
        ver = self._param_version()
        if self._packed is None or ver != self._packed_version:
            self._invalidate()
            self._pack()
            self._packed_version = ver
        B, F, Cc, H, W = x.shape
        if Cc != 3:
            raise NotImplementedError('3-channel frames only')
        key = (B, F, H, W, str(x.device), per_frame_t is not None, self._sdt, self.micro_batches, self.fuse_norm,
               self.pipeline_norm, self.micro_batch_join_hw, self.fused_temporal, self.fuse_head_norm, self.qkv_from_stream)
        ws = self._workspaces.get(key)
        if ws is None:
            chans = sum((n['cin'] + 3 * n['cout']) if n['kind'] == 'res' else 3 * n.get('C', self.model_channels)
                        for n in self.plan) + 2 * self.model_channels
            ws = self._workspaces[key] = self._Workspace(B, F, H, W, x.device, 2 * B * F * chans)
        def plain(v, n, dtype=torch.float32):
            return torch.is_tensor(v) and v.is_cuda and v.dtype == dtype and v.is_contiguous() and v.numel() == n
        if (self.stage_inputs_fused and plain(x, ws.x.numel()) and plain(x0, ws.x.numel()) and x.numel() % 4 == 0
                and plain(obs, B * F) and plain(lat, B * F) and plain(kinda, B * F) and plain(t, B)):
            # one launch instead of seven device-to-device copies around every graph replay
            fi_ok = plain(frame_indices, B * F, torch.long)
            ops.stage_inputs(x, x0, obs, lat, kinda, t, frame_indices if fi_ok else None, ws)
            if not fi_ok:
                ws.fi.copy_(frame_indices.reshape(B, F))
        else:
            ws.x.copy_(x)
            ws.x0.copy_(x0)
            ws.obs.copy_(obs.reshape(B, F))
            ws.lat.copy_(lat.reshape(B, F))
            ws.kinda.copy_(kinda.reshape(B, F))
            ws.t.copy_(t.reshape(B))
            ws.fi.copy_(frame_indices.reshape(B, F))
        if self.use_frame_encoding:
            fi = frame_indices.reshape(B, F).float()
            ws.fi_float.copy_(fi - fi.mean(dim=1, keepdim=True) if self.enforce_position_invariance else fi)
        if per_frame_t is not None:
            ws.t_override.copy_(per_frame_t.reshape(B * F))
        if attn_log is not None:             # attention-map logging: eager launches, per-block RPE tables
            out = self._run(ws, T_attn, per_frame_t is not None, attn_log)
        elif self.use_cuda_graph:
            if ws.graph is None:
                self._run(ws, T_attn, per_frame_t is not None)           # warm-up: allocates every buffer
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run(ws, T_attn, per_frame_t is not None)
                ws.graph = g
            ws.graph.replay()
            out = ws.out
        else:
            out = self._run(ws, T_attn, per_frame_t is not None)
        # gaussian_diffusion._eps consumes eps in the next kernel of the same stream: it borrows the workspace tensor
        # (one flag, cleared by the call it applies to) instead of paying for a copy per step
        borrow, self._borrow_output = self._borrow_output, False
        return out.clone() if (clone and not borrow) else out

    def forward(self, x, timesteps, y=None, attn_mask=None, T=1, return_attn_weights=False, frame_indices=None,
                **kwargs):
        raise NotImplementedError('image-model entry point: use UNetVideoModel / CondMargVideoModel')


class UNetVideoModel(UNetModel):
    def __init__(self, T, use_frame_encoding, cross_frame_attention, enforce_position_invariance, *args, **kwargs):
        if not cross_frame_attention:
            raise NotImplementedError('cross_frame_attention=False is not supported yet')
        self.T = T
        self.use_frame_encoding, self.cross_frame_attention = use_frame_encoding, cross_frame_attention
        self.enforce_position_invariance = enforce_position_invariance
        super().__init__(*args, **kwargs)

    def forward(self, x, timesteps, frame_indices=None, attn_mask=None, return_attn_weights=False, **kwargs):
        """Unconditioned video forward (unet.py:898-912): timesteps holds one value per frame."""
        attn_log = {'spatial': [], 'temporal': [], 'mixed': []} if return_attn_weights else None   # unet.py:797-801
        B, F = x.shape[:2]
        if frame_indices is None:
            frame_indices = torch.arange(F, device=x.device).view(1, F).expand(B, F)
        ones = torch.ones(B, F, device=x.device)
        zeros = torch.zeros(B, F, device=x.device)
        mask = ones if attn_mask is None else attn_mask.reshape(B, F).float()
        # latent everywhere -> the conditioning mix is the identity; attention mask = given mask
        out = self._execute(x, x, zeros, mask, zeros, timesteps.reshape(-1)[:B].float(), frame_indices, F,
                            per_frame_t=timesteps.reshape(B * F).float(), attn_log=attn_log)
        return out, attn_log


class CondMargVideoModel(UNetVideoModel):
    def __init__(self, cond_emb_type, **kwargs):
        """unet.py:932-947: 'channel' adds two indicator channels, 'duplicate' / 'all' double the input channels,
        't=0' marks observed frames through the timestep only."""
        base = cond_emb_type.replace('-initzero', '')
        if 'channel' in cond_emb_type:
            kwargs['in_channels'] += 2
        elif 'duplicate' in cond_emb_type or 'all' in cond_emb_type:
            kwargs['in_channels'] *= 2
        elif cond_emb_type != 't=0':
            raise NotImplementedError(f'cond_emb_type={cond_emb_type!r}')
        super().__init__(**kwargs)
        w_in = self.input_blocks._modules['0']._modules['0'].weight.data
        if cond_emb_type == 'channel-initzero':
            w_in[:, 3] = 0.0
        if cond_emb_type in ('duplicate-initzero', 'all-initzero'):
            w_in[:, 3:] = w_in[:, :3]
        self.cond_emb_type = base
        if base not in ('channel', 'duplicate', 'all', 't=0'):
            raise NotImplementedError(f'cond_emb_type={cond_emb_type!r}')
        self._cond_mode = {'channel': 0, 'duplicate': 1, 'all': 1, 't=0': 2}[base]

    def forward(self, x, x0=None, obs_mask=None, latent_mask=None, kinda_marg_mask=None, timesteps=None,
                frame_indices=None, return_attn_weights=False, **kwargs):
        """unet.py:949-1026.  Called as model(x, t, **model_kwargs) (gaussian_diffusion.py:274) or
        model(x, timesteps=t, **model_kwargs) (respace.py:119)."""
        if timesteps is None:
            raise TypeError('timesteps is required')
        attn_log = {'spatial': [], 'temporal': [], 'mixed': []} if return_attn_weights else None   # unet.py:797-801
        B, F = x.shape[:2]
        if frame_indices is None:
            frame_indices = torch.arange(F, device=x.device).view(1, F).expand(B, F)
        t = timesteps.float()
        observed, per_frame_t = x0, None
        if self.cond_emb_type == 'channel':
            if 'x_t_minus_1' not in kwargs or 'observed_frames' not in kwargs:
                raise KeyError('x_t_minus_1')   # the reference requires both kwargs (unet.py:958-974)
            which = kwargs['observed_frames']
            om = obs_mask.reshape(B, F).float()
            if 'hybrid' in which:
                # unet.py:966-974, 1001-1009: below the threshold the observed frames are x_{t-1} at time t-1,
                # above it kwargs['hybrid'] at time `threshold`
                thr = int(which.split('_')[-1])
                below = (t < thr).view(B, 1, 1, 1, 1)
                observed = torch.where(below, kwargs['x_t_minus_1'], kwargs['hybrid'])
                t_obs = torch.where(t < thr, t - 1, torch.full_like(t, float(thr)))
            elif which == 'x_0':
                t_obs = None                    # observed frames at time 0: the kernel's own per-frame timestep
            elif which == 'x_t':
                observed, t_obs = x, t
            elif which == 'x_t_minus_1':
                observed, t_obs = kwargs['x_t_minus_1'], t - 1
            else:
                raise NotImplementedError(f'observed_frames={which!r} (training-only or unknown)')
            if t_obs is not None:
                per_frame_t = t_obs.view(B, 1) * om + t.view(B, 1) * (1 - om)
        elif self.cond_emb_type == 't=0':
            # unet.py:1019 writes -1 through an expanded (stride-0) view of the timesteps: every frame of a video
            # that has at least one observed frame ends up at t = -1.  Reproduced as the reference behaves.
            any_obs = (obs_mask.reshape(B, F) == 1).any(dim=1)
            per_frame_t = torch.where(any_obs, torch.full_like(t, -1.0), t).view(B, 1).expand(B, F).contiguous()
        out = self._execute(x, observed, obs_mask, latent_mask, kinda_marg_mask, t, frame_indices, F,
                            per_frame_t=per_frame_t, attn_log=attn_log)
        return out, attn_log

    def __call__(self, x, *args, **kwargs):
        # model(x, t, x0=..., ...) : map the positional timestep onto the keyword
        if args:
            kwargs['timesteps'] = args[0]
            args = args[1:]
        return super().__call__(x, *args, **kwargs)


def param_spec(model):
    return {k: list(v.shape) for k, v in model.state_dict().items()}
