"""CPU oracle for the video U-Net denoising forward.  TEST INFRASTRUCTURE ONLY.

This is a from-scratch functional restatement (plain torch fp32 on the CPU) of
the algorithm in the reference `improved_diffusion/unet.py`.  It works directly
on a reference-format ``state_dict`` (flat ``{key: tensor}``), has no nn.Module
and shares no code with the product package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import it.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so this file is pinned against outputs of the reference itself,
generated in the build container by ``oracle/make_golden.py`` and committed as
``tests/golden/*.npz`` (see ``tests/test_oracle_golden.py``).

Reference citations (file:line into the reference checkout):
  conditioning mix ............ improved_diffusion/unet.py:949-1026
  (B,F,..)->(B*F,..) wrapper ... improved_diffusion/unet.py:898-926
  U-Net wiring ................. improved_diffusion/unet.py:564-749, 768-839
  ResBlock ..................... improved_diffusion/unet.py:118-198
  Up/Downsample ................ improved_diffusion/unet.py:47-101
  factorized attention ......... improved_diffusion/unet.py:236-268
  RPE attention ................ improved_diffusion/unet.py:471-540
  RPE net / einsums ............ improved_diffusion/unet.py:283-298, 357-378
  GroupNorm32/SiLU/t-embedding . improved_diffusion/nn.py:10-17, 89-107
"""
import math

import torch
import torch.nn.functional as F

CHANNEL_MULT = {256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4), 32: (1, 2, 2, 2)}


def model_config(image_size, num_channels=128, num_res_blocks=2, num_heads=4,
                 attention_resolutions='16,8', use_scale_shift_norm=True,
                 use_spatial_encoding=True, allow_interactions_between_padding=True,
                 learn_sigma=False, use_rpe_net=True, rp_alpha=None, rp_beta=None, rp_gamma=None,
                 use_frame_encoding=False, enforce_position_invariance=False, T=None, cond_emb_type='channel',
                 **_ignored):
    """Architecture facts the oracle needs (script_util.py:229-300)."""
    return dict(
        rpe_net=use_rpe_net, rp=(rp_alpha, rp_beta, rp_gamma),
        frame_enc=use_frame_encoding, pos_inv=enforce_position_invariance, T=T,
        cond=cond_emb_type.replace('-initzero', ''),
        image_size=image_size, ch=num_channels, nrb=num_res_blocks, heads=num_heads,
        mult=CHANNEL_MULT[image_size],
        attn_ds=tuple(image_size // int(r) for r in attention_resolutions.split(',')),
        scale_shift=use_scale_shift_norm, spatial_enc=use_spatial_encoding,
        pad_interact=allow_interactions_between_padding,
        out_ch=6 if learn_sigma else 3,
    )


def block_plan(cfg):
    """Enumerate modules the way the reference constructor does (unet.py:616-742).

    Returns (input_plan, middle_plan, output_plan, n_blocks_before_attn); each plan
    is a list (one entry per `*_blocks.i`) of lists of (kind, sub_index) with kind
    in {'conv', 'res', 'attn', 'down', 'up'}.
    """
    ch_mult, nrb, attn_ds = cfg['mult'], cfg['nrb'], cfg['attn_ds']
    inp = [[('conv', 0)]]
    ds = 1
    before_attn = None
    for level in range(len(ch_mult)):
        for _ in range(nrb):
            if ds in attn_ds and before_attn is None:
                before_attn = len(inp)
            mods = [('res', 0)]
            if ds in attn_ds:
                mods.append(('attn', 1))
            inp.append(mods)
        if level != len(ch_mult) - 1:
            inp.append([('down', 0)])
            ds *= 2
    if before_attn is None:
        before_attn = len(inp)
    mid = [('res', 0), ('attn', 1), ('res', 2)]
    out = []
    for level in reversed(range(len(ch_mult))):
        for i in range(nrb + 1):
            mods = [('res', 0)]
            if ds in attn_ds:
                mods.append(('attn', 1))
            if level and i == nrb:
                mods.append(('up', len(mods)))
                ds //= 2
            out.append(mods)
    return inp, mid, out, before_attn


def silu(x):
    return x * torch.sigmoid(x)


def gn32(x, w, b):
    """GroupNorm32: 32 groups, eps 1e-5, computed in fp32 (nn.py:15-17)."""
    return F.group_norm(x.float(), 32, w, b, eps=1e-5).type(x.dtype)


def sinusoid(t, dim, max_period=10000):
    """[cos | sin] embedding (nn.py:89-107)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def res_block(sd, p, x, emb, scale_shift):
    """unet.py:185-198."""
    h = F.conv2d(silu(gn32(x, sd[p + 'in_layers.0.weight'], sd[p + 'in_layers.0.bias'])),
                 sd[p + 'in_layers.2.weight'], sd[p + 'in_layers.2.bias'], padding=1)
    e = F.linear(silu(emb), sd[p + 'emb_layers.1.weight'], sd[p + 'emb_layers.1.bias'])[:, :, None, None]
    gw, gb = sd[p + 'out_layers.0.weight'], sd[p + 'out_layers.0.bias']
    if scale_shift:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = gn32(h, gw, gb) * (1 + scale) + shift
        h = silu(h)
    else:
        h = silu(gn32(h + e, gw, gb))
    h = F.conv2d(h, sd[p + 'out_layers.3.weight'], sd[p + 'out_layers.3.bias'], padding=1)
    if p + 'skip_connection.weight' in sd:
        w = sd[p + 'skip_connection.weight']
        x = F.conv2d(x, w, sd[p + 'skip_connection.bias'], padding=w.shape[-1] // 2)
    return x + h


def rpe_table(sd, p, temb, dist, heads):
    """R[b,i,j,h,f] from the RPE net (unet.py:283-298).  temb (B,T,E), dist (B,T,T) int."""
    d = dist.float()
    feats = torch.stack([torch.log1p(d.clamp(min=0)), torch.log1p((-d).clamp(min=0)), (dist == 0).float()], dim=-1)
    B, T, _ = dist.shape
    e_t = F.linear(temb, sd[p + 'embed_diffusion_time.weight'], sd[p + 'embed_diffusion_time.bias'])
    e_d = F.linear(feats, sd[p + 'embed_distances.weight'], sd[p + 'embed_distances.bias'])
    r = F.linear(silu(e_t[:, :, None, :] + e_d), sd[p + 'out.weight'], sd[p + 'out.bias'])
    C = r.shape[-1]
    return r.view(B, T, T, heads, C // heads)


def rpe_bucket_ids(dist, alpha, beta, gamma):
    """Piecewise index function of the lookup-table RPE (unet.py:326-338, eq. 18 of arXiv 2107.14222): identity up to
    alpha, logarithmic up to gamma, clipped at beta; float32 arithmetic and truncation as the reference does them."""
    ids = dist.clone()
    far = ids.abs() > alpha
    if far.any():
        coef = torch.log(ids[far].abs() / alpha) / math.log(gamma / alpha)
        mag = torch.minimum(torch.tensor(beta), alpha + coef * (beta - alpha)).int()
        ids[far] = mag * torch.sign(ids[far])
    return ids


def rpe_lookup(sd, p, dist, rp):
    """R[b,i,j,h,f] = table[bucket(dist)] (unet.py:340-347); negative buckets index from the end of the table."""
    return sd[p + 'lookup_table_weight'][rpe_bucket_ids(dist, *rp)]


def rpe_attention(sd, p, x, temb, frame_indices, attn_mask, heads, pad_interact, with_rpe, rp=None, log=None):
    """unet.py:471-540.  x: (B, D, C, L) attends over the last axis L."""
    B, D, C, L = x.shape
    hd = C // heads
    scale = hd ** -0.5
    xn = gn32(x.reshape(B * D, C, L), sd[p + 'norm.weight'], sd[p + 'norm.bias']).view(B, D, C, L)
    xn = xn.permute(0, 1, 3, 2)                                   # B D L C
    qkv = F.linear(xn, sd[p + 'qkv.weight'], sd[p + 'qkv.bias']).reshape(B, D, L, 3, heads, hd)
    q, k, v = (qkv[:, :, :, i].permute(0, 1, 3, 2, 4) for i in range(3))   # B D H L hd
    q = q * scale
    logits = q @ k.transpose(-1, -2)                              # B D H L L
    if with_rpe:
        dist = frame_indices[:, :, None] - frame_indices[:, None, :]
        tb = temb.view(B, L, -1)
        if p + 'rpe_k.lookup_table_weight' in sd:          # use_rpe_net=False
            r_k, r_q, r_v = (rpe_lookup(sd, p + f'rpe_{n}.', dist, rp) for n in 'kqv')
        else:
            r_k = rpe_table(sd, p + 'rpe_k.rpe_net.', tb, dist, heads)
            r_q = rpe_table(sd, p + 'rpe_q.rpe_net.', tb, dist, heads)
            r_v = rpe_table(sd, p + 'rpe_v.rpe_net.', tb, dist, heads)
        logits = logits + torch.einsum('bdhtf,btshf->bdhts', q, r_k)
        logits = logits + torch.einsum('bdhtf,btshf->bdhts', k * scale, r_q).transpose(-1, -2)
    if attn_mask is not None:
        m = attn_mask.view(B, L).float()
        allowed = m[:, None, :] * m[:, :, None]
        if pad_interact:
            allowed = allowed + (1 - m[:, None, :]) * (1 - m[:, :, None])
        else:
            idx = torch.arange(L)
            allowed[:, idx, idx] = 1.0
        neg = torch.zeros_like(allowed)
        neg[allowed == 0] = float('inf')
        logits = logits - neg.view(B, 1, 1, L, L)
    w = torch.softmax(logits.float(), dim=-1)
    if log is not None:                                           # logging maps, unet.py:464-468
        log.append(w.reshape(B * D, heads, L, L).mean(dim=1).abs())
    out = w @ v
    if with_rpe:
        out = out + torch.einsum('bdhts,btshf->bdhtf', w, r_v)
    out = out.permute(0, 1, 3, 2, 4).reshape(B, D, L, C)
    out = F.linear(out, sd[p + 'proj_out.weight'], sd[p + 'proj_out.bias'])
    return (xn + out).permute(0, 1, 3, 2)                          # residual on the NORMALISED x


def factorized_attention(sd, p, x, temb, frame_indices, attn_mask, T, cfg, attn_log=None):
    """unet.py:236-268: temporal RPE attention, then spatial attention."""
    BT, C, H, W = x.shape
    B = BT // T
    xt = x.view(B, T, C, H, W).permute(0, 3, 4, 2, 1).reshape(B, H * W, C, T)
    xt = rpe_attention(sd, p + 'temporal_attention.', xt, temb, frame_indices,
                       attn_mask.reshape(B, T), cfg['heads'], cfg['pad_interact'], True, rp=cfg.get('rp'),
                       log=None if attn_log is None else attn_log['temporal'])
    xs = xt.view(B, H, W, C, T).permute(0, 4, 3, 1, 2).reshape(B, T, C, H * W)
    xs = rpe_attention(sd, p + 'spatial_attention.', xs, temb, None, None,
                       cfg['heads'], cfg['pad_interact'], False,
                       log=None if attn_log is None else attn_log['spatial'])
    return xs.reshape(BT, C, H, W)


def unet_forward(sd, cfg, x, timesteps, frame_indices, attn_mask, T, taps=None, attn_log=None):
    """UNetModel.forward (unet.py:768-839) on x (B*T, Cin, H, W), timesteps (B*T,)."""
    inp, mid, outp, before_attn = block_plan(cfg)
    emb = F.linear(sinusoid(timesteps, cfg['ch']), sd['time_embed.0.weight'], sd['time_embed.0.bias'])
    emb = F.linear(silu(emb), sd['time_embed.2.weight'], sd['time_embed.2.bias'])
    if taps is not None:
        taps['emb'] = emb

    def run(prefix, mods, h):
        for kind, j in mods:
            p = f'{prefix}{j}.'
            if kind == 'conv':
                h = F.conv2d(h, sd[p + 'weight'], sd[p + 'bias'], padding=1)
            elif kind == 'res':
                h = res_block(sd, p, h, emb, cfg['scale_shift'])
            elif kind == 'attn':
                h = factorized_attention(sd, p, h, emb, frame_indices, attn_mask, T, cfg, attn_log)
            elif kind == 'down':
                h = F.conv2d(h, sd[p + 'op.weight'], sd[p + 'op.bias'], stride=2, padding=1)
            elif kind == 'up':
                h = F.interpolate(h, scale_factor=2, mode='nearest')
                h = F.conv2d(h, sd[p + 'conv.weight'], sd[p + 'conv.bias'], padding=1)
            if taps is not None:
                taps[p + kind] = h
        return h

    h = x.float()
    hs = []
    for i, mods in enumerate(inp):
        h = run(f'input_blocks.{i}.', mods, h)
        hs.append(h)
        if i + 1 == before_attn:                                  # unet.py:816-818, 841-844, 914-926
            if cfg['spatial_enc']:
                h = h + sd['spatial_encoding']
            if cfg.get('frame_enc'):
                fi = frame_indices.float()
                if cfg['pos_inv']:
                    fi = fi - fi.mean(dim=1, keepdim=True)
                h = h + sinusoid(fi.reshape(-1), h.shape[1], max_period=cfg['T'] * 10)[:, :, None, None]
    h = run('middle_block.', mid, h)
    for i, mods in enumerate(outp):
        h = run(f'output_blocks.{i}.', mods, torch.cat([h, hs.pop()], dim=1))
    h = silu(gn32(h, sd['out.0.weight'], sd['out.0.bias']))
    return F.conv2d(h, sd['out.2.weight'], sd['out.2.bias'], padding=1)


def video_forward(sd, cfg, x, timesteps, frame_indices=None, attn_mask=None):
    """UNetVideoModel.forward (unet.py:898-912): x (B,F,3,H,W), timesteps (B,F) -- one value per frame --,
    attn_mask (B,F,1,1,1) (the reference flattens it, so it is required), no conditioning mix."""
    B, Fr, C, H, W = x.shape
    if frame_indices is None:
        frame_indices = torch.arange(Fr).view(1, Fr).expand(B, Fr)
    out = unet_forward(sd, cfg, x.reshape(B * Fr, C, H, W), timesteps.reshape(B * Fr), frame_indices, attn_mask, Fr)
    return out.view(B, Fr, cfg['out_ch'], H, W)


def cond_marg_forward(sd, cfg, x, x0, obs_mask, latent_mask, kinda_marg_mask, timesteps,
                      frame_indices=None, taps=None, observed_frames='x_0', x_t_minus_1=None, hybrid=None,
                      attn_log=None):
    """CondMargVideoModel.forward in eval mode (unet.py:949-1026 + 898-912) for cond_emb_type 'channel' (every
    inference-time observed_frames choice), 'duplicate' / 'all' and 't=0'.
    x, x0: (B,F,3,H,W); masks (B,F,1,1,1); timesteps (B,)."""
    B, Fr, C, H, W = x.shape
    anything = (obs_mask + latent_mask + kinda_marg_mask).clamp(max=1)
    t = timesteps.view(B, 1).expand(B, Fr)
    om = obs_mask.view(B, Fr)
    cond = cfg.get('cond', 'channel')
    if cond == 'channel':
        if 'hybrid' in observed_frames:
            thr = int(observed_frames.split('_')[-1])
            below = (t < thr).int()
            b5 = below[:, :, None, None, None]
            observed = x_t_minus_1 * b5 + hybrid * (1 - b5)
            t_obs = below * (t - 1) + (1 - below) * torch.ones_like(t) * thr
        else:
            observed = {'x_0': x0, 'x_t': x, 'x_t_minus_1': x_t_minus_1}[observed_frames]
            t_obs = {'x_0': torch.zeros_like(t), 'x_t': t, 'x_t_minus_1': t - 1}[observed_frames]
        ones = torch.ones_like(x[:, :, :1])
        x_in = torch.cat([x * latent_mask + observed * obs_mask + x * (1 - anything),
                          ones * obs_mask, ones * kinda_marg_mask], dim=2)
        t = t_obs * om + t * (1 - om)
    elif cond in ('duplicate', 'all'):
        x_in = torch.cat([x * latent_mask + x * (1 - anything), x0 * obs_mask], dim=2)
    elif cond == 't=0':
        # the reference assigns -1 through a stride-0 expanded view (unet.py:1019): one observed frame sends the whole
        # video's timesteps to -1
        x_in = x
        t = torch.where((om == 1).any(dim=1, keepdim=True), torch.full_like(t, -1.0), t)
    else:
        raise ValueError(cond)
    if frame_indices is None:
        frame_indices = torch.arange(Fr).view(1, Fr).expand(B, Fr)
    out = unet_forward(sd, cfg, x_in.reshape(B * Fr, x_in.shape[2], H, W), t.reshape(B * Fr),
                       frame_indices, anything, Fr, taps=taps, attn_log=attn_log)
    return out.view(B, Fr, cfg['out_ch'], H, W)
