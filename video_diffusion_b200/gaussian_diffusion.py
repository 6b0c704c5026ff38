"""Diffusion sampling / ELBO driver with the reference's API, computing on libvdm kernels.

Mirrors `improved_diffusion/gaussian_diffusion.py` of the reference for the inference path:
same constructor, attributes (float64 numpy tables) and method signatures / return dict keys
(`q_sample` :190, `p_mean_variance` :229, `p_sample` :403, `p_sample_loop[_progressive]`
:450/:528, `ddim_sample` :597, `ddim_sample_loop[_progressive]` :670/:702, `_vb_terms_bpd`
:750, `_prior_bpd` :909, `calc_bpd_loop[_subsampled]` :928/:1004, `q_posterior_mean_variance` :208,
`_predict_xstart_from_eps / _from_xprev / _predict_eps_from_xstart` :374-396, `ddim_reverse_sample`
:636).  Everything after the
network call is ONE fused elementwise kernel per step (`vdm_sampler_step`, `vdm_vb_terms`),
with the schedule coefficients gathered on the device from tables uploaded once -- the
reference re-uploads eight 1000-entry float64 tables per step (:1019-1031).

Unsupported on this path (raise NotImplementedError, no silent fallback): learned variances,
x_{t-1}/x_0-predicting models, `use_gradient_method`, `denoised_fn`, `training_losses`.
"""
import enum
import math

import numpy as np
import torch as th

from . import _lib, ops


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """Same schedules as the reference (gaussian_diffusion.py:20-50)."""
    n = num_diffusion_timesteps
    if schedule_name in ('linear', 'noisier_linear'):
        scale = 1000 / n
        end = 0.02 if schedule_name == 'linear' else 0.025
        return np.linspace(scale * 0.0001, scale * end, n, dtype=np.float64)
    if schedule_name == 'cosine':
        return betas_for_alpha_bar(n, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f'unknown beta schedule: {schedule_name}')


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


def device_tables(diffusion=None, s_like=None, var_type=ModelVarType.FIXED_LARGE):
    """[TAB_COUNT][T] fp32 coefficient table consumed by the sampler / ELBO kernels.  `s_like`
    lets tests build it from any object with the oracle's attribute names."""
    if s_like is not None:
        src = dict(SQRT_RECIP_ACP=s_like.sqrt_recip_acp, SQRT_RECIPM1_ACP=s_like.sqrt_recipm1_acp,
                   POST_C1=s_like.post_c1, POST_C2=s_like.post_c2, MODEL_LOGVAR=s_like.model_logvar,
                   MODEL_VAR=s_like.model_var, ACP=s_like.acp, ACP_PREV=s_like.acp_prev,
                   POST_LOGVAR=s_like.post_logvar, SQRT_ACP=s_like.sqrt_acp, SQRT_1M_ACP=s_like.sqrt_1m_acp,
                   LOG_1M_ACP=s_like.log_1m_acp, POST_VAR=s_like.post_var, RECIP_POST_C1=s_like.recip_post_c1,
                   POST_C2_DIV_C1=s_like.post_c2_div_c1, ACP_NEXT=s_like.acp_next)
    else:
        d = diffusion
        if var_type == ModelVarType.FIXED_LARGE:     # gaussian_diffusion.py:300-309
            mv = np.append(d.posterior_variance[1], d.betas[1:])
            mlv = np.log(mv)
        else:
            mv, mlv = d.posterior_variance, d.posterior_log_variance_clipped
        src = dict(SQRT_RECIP_ACP=d.sqrt_recip_alphas_cumprod, SQRT_RECIPM1_ACP=d.sqrt_recipm1_alphas_cumprod,
                   POST_C1=d.posterior_mean_coef1, POST_C2=d.posterior_mean_coef2, MODEL_LOGVAR=mlv, MODEL_VAR=mv,
                   ACP=d.alphas_cumprod, ACP_PREV=d.alphas_cumprod_prev, POST_LOGVAR=d.posterior_log_variance_clipped,
                   SQRT_ACP=d.sqrt_alphas_cumprod, SQRT_1M_ACP=d.sqrt_one_minus_alphas_cumprod,
                   LOG_1M_ACP=d.log_one_minus_alphas_cumprod, POST_VAR=d.posterior_variance,
                   RECIP_POST_C1=1.0 / d.posterior_mean_coef1,
                   POST_C2_DIV_C1=d.posterior_mean_coef2 / d.posterior_mean_coef1, ACP_NEXT=d.alphas_cumprod_next)
    n = len(src['ACP'])
    tab = np.zeros((_lib.TAB_COUNT, n), dtype=np.float64)
    for name, row in _lib.TAB.items():
        tab[row] = src[name]
    return th.from_numpy(tab).float()      # the reference casts table[t] with .float() (:1028)


class GaussianDiffusion:
    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas = alphas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self._tables = {}

    # ---- device-side state -------------------------------------------------------------
    def tables(self, device):
        key = str(device)
        if key not in self._tables:
            if self.model_var_type not in (ModelVarType.FIXED_LARGE, ModelVarType.FIXED_SMALL):
                raise NotImplementedError('learned variances are not supported by the CUDA sampler kernels')
            self._tables[key] = device_tables(self, var_type=self.model_var_type).to(device).contiguous()
        return self._tables[key]

    def _require_eps_model(self):
        if self.model_mean_type != ModelMeanType.EPSILON:
            raise NotImplementedError('only epsilon-predicting models are supported by the CUDA sampler kernels')

    def _row(self, name, t, like):
        v = self.tables(like.device)[_lib.TAB[name]][t]
        return v.view(-1, *([1] * (like.dim() - 1))).expand(like.shape)

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def _eps(self, model, x, t, model_kwargs, return_attn_weights=False):
        # every caller feeds eps to a libvdm kernel before the next forward: let the B200 model hand out its workspace
        # tensor instead of a copy (a stand-in network has no such flag and is called as is)
        inner = getattr(model, 'model', model)
        borrow = hasattr(inner, '_borrow_output')
        if borrow:
            inner._borrow_output = True
        try:
            out, attn = model(x, self._scale_timesteps(t), return_attn_weights=return_attn_weights,
                              **(model_kwargs or {}))
        finally:
            if borrow:
                inner._borrow_output = False
        if out.shape != x.shape:
            raise NotImplementedError('model output shape %s != input shape %s (learned sigma?)'
                                      % (tuple(out.shape), tuple(x.shape)))
        return out.contiguous(), attn

    # ---- forward process ---------------------------------------------------------------
    def _lincomb(self, op, a, b, t, row_a, row_b='ACP'):
        a = a.contiguous()
        with th.cuda.device(a.device):
            return ops.lincomb(op, a, None if b is None else b.contiguous(), t.long().contiguous(),
                               self.tables(a.device), _lib.TAB[row_a], _lib.TAB[row_b])

    def q_mean_variance(self, x_start, t):
        """q(x_t | x_0) (gaussian_diffusion.py:171-188)."""
        return (self._lincomb(3, x_start, None, t, 'SQRT_ACP'), 1.0 - self._row('ACP', t, x_start),
                self._row('LOG_1M_ACP', t, x_start))

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        t = t.long()
        t = th.where(t < 0, t + self.num_timesteps, t)     # numpy-style negative index (the loop asks for t - 1 at t = 0)
        with th.cuda.device(x_start.device):
            return ops.q_sample(x_start.contiguous(), noise.contiguous(), t.contiguous(),
                                self.tables(x_start.device))

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """q(x_{t-1} | x_t, x_0) (gaussian_diffusion.py:208-227)."""
        assert x_start.shape == x_t.shape
        mean = self._lincomb(0, x_start, x_t, t, 'POST_C1', 'POST_C2')
        return mean, self._row('POST_VAR', t, x_t), self._row('POST_LOGVAR', t, x_t)

    # ---- reverse process ---------------------------------------------------------------
    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                        return_attn_weights=False, use_gradient_method=False):
        if use_gradient_method or denoised_fn is not None:
            raise NotImplementedError('use_gradient_method / denoised_fn need autograd or host callbacks')
        self._require_eps_model()
        assert t.shape == (x.shape[0],)
        x = x.contiguous()
        eps, attn = self._eps(model, x, t, model_kwargs, return_attn_weights)
        pred, mean, scratch = th.empty_like(x), th.empty_like(x), th.empty_like(x)
        with th.cuda.device(x.device):
            ops.sampler_step(0, x, eps, x, t.long().contiguous(), self.tables(x.device), clip_denoised=clip_denoised,
                             sample=scratch, pred_xstart=pred, mean=mean)
        return {'mean': mean, 'variance': self._row('MODEL_VAR', t, x), 'log_variance': self._row('MODEL_LOGVAR', t, x),
                'pred_xstart': pred, 'attn': attn}

    def _predict_xstart_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return self._lincomb(1, x_t, eps, t, 'SQRT_RECIP_ACP', 'SQRT_RECIPM1_ACP')

    def _predict_xstart_from_xprev(self, x_t, t, xprev):
        """(xprev - coef2 * x_t) / coef1 in the reference's form (gaussian_diffusion.py:384-390)."""
        assert x_t.shape == xprev.shape
        return self._lincomb(1, xprev, x_t, t, 'RECIP_POST_C1', 'POST_C2_DIV_C1')

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return self._lincomb(2, x_t, pred_xstart, t, 'SQRT_RECIP_ACP', 'SQRT_RECIPM1_ACP')

    def _step(self, mode, model, x, t, clip_denoised, model_kwargs, eta=0.0, return_attn_weights=False):
        self._require_eps_model()
        x = x.contiguous()
        eps, attn = self._eps(model, x, t, model_kwargs, return_attn_weights)
        # one draw per step, same order as the reference (:438, :628); the reverse ODE draws nothing (:636-668)
        noise = th.randn_like(x) if mode != 2 else x
        pred = th.empty_like(x)
        with th.cuda.device(x.device):
            sample = ops.sampler_step(mode, x, eps, noise, t.long().contiguous(), self.tables(x.device),
                                      clip_denoised=clip_denoised, eta=eta, pred_xstart=pred)
        return sample, pred, attn

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                 return_attn_weights=False, use_gradient_method=False):
        if use_gradient_method or denoised_fn is not None:
            raise NotImplementedError('use_gradient_method / denoised_fn need autograd or host callbacks')
        sample, pred, attn = self._step(0, model, x, t, clip_denoised, model_kwargs,
                                        return_attn_weights=return_attn_weights)
        return {'sample': sample, 'pred_xstart': pred, 'attn': attn}

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0):
        if denoised_fn is not None:
            raise NotImplementedError('denoised_fn is a host callback')
        sample, pred, _ = self._step(1, model, x, t, clip_denoised, model_kwargs, eta=eta)
        return {'sample': sample, 'pred_xstart': pred}

    def ddim_reverse_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, eta=0.0):
        """x_{t+1} from x_t with the DDIM reverse ODE (gaussian_diffusion.py:636-668)."""
        assert eta == 0.0, 'Reverse ODE only for deterministic path'
        if denoised_fn is not None:
            raise NotImplementedError('denoised_fn is a host callback')
        sample, pred, _ = self._step(2, model, x, t, clip_denoised, model_kwargs)
        return {'sample': sample, 'pred_xstart': pred}

    def _loop(self, step_fn, model, shape, noise, device, progress):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = th.full((shape[0],), i, device=device, dtype=th.long)
            with th.no_grad():
                out = step_fn(img, t)
            yield out
            img = out['sample']
        ops.check_timesteps()

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                  model_kwargs=None, latent_mask=None, device=None, progress=False,
                                  return_attn_weights=False, use_gradient_method=False):
        """Ancestral chain from noise (gaussian_diffusion.py:528-595).  Every step re-noises x0 into the
        x_t_minus_1 (and, for 'hybrid_<N>', hybrid) conditioning tensors with the reference's noise choice: the
        chain's initial `noise` when one was given, fresh draws otherwise.  The draw for the training-only x_random
        is kept (and discarded) so a seeded torch generator stays in step with the reference.  Works on a copy of
        model_kwargs: the caller's dict is left untouched, and no `.cuda()` is hard-coded."""
        kw = dict(model_kwargs or {})
        kw.setdefault('observed_frames', 'x_0')
        which = kw['observed_frames']
        x0 = kw.get('x0')
        if x0 is None:
            raise KeyError('x0')
        kw.setdefault('x_t_minus_1', x0)

        def step(img, t):
            draw = (lambda: th.randn_like(x0)) if noise is None else (lambda: noise)
            n_prev, n_random = draw(), draw()
            del n_random                              # x_random only reaches the network in training mode
            if which != 'x_0':
                kw['x_t_minus_1'] = self.q_sample(x0, t - 1, noise=n_prev)
            if 'hybrid' in which:
                thr = int(which.split('_')[-1])
                if thr >= self.num_timesteps:        # the reference indexes its schedule tables with it
                    raise IndexError(f'hybrid threshold {thr} is out of range for {self.num_timesteps} timesteps')
                kw['hybrid'] = self.q_sample(x0, th.full_like(t, thr), noise=draw())
            return self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                 model_kwargs=kw, return_attn_weights=return_attn_weights,
                                 use_gradient_method=use_gradient_method)

        return self._loop(step, model, shape, noise, device, progress)

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                      latent_mask=None, device=None, progress=False, return_attn_weights=False,
                      use_gradient_method=False):
        """Returns (sample, attns).  With return_attn_weights the per-layer attention maps are averaged per quartile
        of the chain under the keys 'attn/q<quartile>-<temporal|spatial>' (gaussian_diffusion.py:496-524): batch-mean
        over the non-attended axis, spatial maps resized (nearest) to the first layer's size and renormalised."""
        final, attns = None, {}
        chain = self.p_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                               denoised_fn=denoised_fn, model_kwargs=model_kwargs,
                                               latent_mask=latent_mask, device=device, progress=progress,
                                               return_attn_weights=return_attn_weights,
                                               use_gradient_method=use_gradient_method)
        for neg_t, final in enumerate(chain):
            if not return_attn_weights:
                continue
            quartile = (4 * (self.num_timesteps - neg_t - 1)) // self.num_timesteps
            for key, layers in final['attn'].items():
                if len(layers) == 0:
                    continue
                tag = f'attn/q{quartile}-{key}'
                largest = layers[0][0].shape
                for layer in layers:
                    B = shape[0]
                    layer = layer.view(B, layer.shape[0] // B, *layer.shape[1:]).mean(dim=1)
                    if 'temporal' not in key:
                        resized = th.nn.functional.interpolate(layer.unsqueeze(0), size=largest, mode='nearest').squeeze(0)
                        layer = resized / resized.mean() * layer.mean()
                    attns[tag] = attns.get(tag, 0) + layer / (self.num_timesteps / 4)
        return final['sample'], attns

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     model_kwargs=None, latent_mask=None, device=None, progress=False, eta=0.0):
        return self._loop(lambda img, t: self.ddim_sample(model, img, t, clip_denoised=clip_denoised,
                                                          denoised_fn=denoised_fn, model_kwargs=model_kwargs, eta=eta),
                          model, shape, noise, device, progress)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, model_kwargs=None,
                         latent_mask=None, device=None, progress=False, eta=0.0):
        final = None
        for final in self.ddim_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                       denoised_fn=denoised_fn, model_kwargs=model_kwargs,
                                                       device=device, progress=progress, eta=eta):
            pass
        return final['sample']

    # ---- ELBO --------------------------------------------------------------------------
    @staticmethod
    def _frame_mask(latent_mask, x):
        B, F = x.shape[0], x.shape[1]
        if latent_mask is None:
            return th.ones(B, F, device=x.device)
        return latent_mask.reshape(B, F).float().contiguous()

    def _vb_terms_bpd(self, model, x_start, x_t, t, clip_denoised=True, model_kwargs=None, latent_mask=None):
        self._require_eps_model()
        x_start, x_t = x_start.contiguous(), x_t.contiguous()
        eps, _ = self._eps(model, x_t, t, model_kwargs)
        tl = t.long().contiguous()
        acc = th.zeros(x_t.shape[0], 3, device=x_t.device, dtype=th.float64)
        pred, scratch = th.empty_like(x_t), th.empty_like(x_t)
        with th.cuda.device(x_t.device):
            ops.vb_terms(x_start, x_t, eps, eps, tl, self.tables(x_t.device), self._frame_mask(latent_mask, x_t),
                         clip_denoised, acc)
            ops.sampler_step(0, x_t, eps, x_t, tl, self.tables(x_t.device), clip_denoised=clip_denoised,
                             sample=scratch, pred_xstart=pred)
        return {'output': acc[:, 0].float(), 'pred_xstart': pred}

    def _prior_bpd(self, x_start, latent_mask=None):
        acc = th.zeros(x_start.shape[0], device=x_start.device, dtype=th.float64)
        with th.cuda.device(x_start.device):
            ops.prior_bpd(x_start.contiguous(), self.tables(x_start.device), self._frame_mask(latent_mask, x_start),
                          acc)
        return acc.float()

    def calc_bpd_loop_subsampled(self, model, x_start, clip_denoised=True, model_kwargs=None, latent_mask=None,
                                 t_seq=None):
        """One q_sample kernel + one U-Net forward + one fused ELBO-terms kernel per timestep; the
        (B, n_t, 3) accumulator stays on the device until the end."""
        self._require_eps_model()
        device = x_start.device
        x_start = x_start.contiguous()
        B = x_start.shape[0]
        if t_seq is None:
            t_seq = list(range(self.num_timesteps))[::-1]
        two_d = isinstance(t_seq, np.ndarray) and t_seq.ndim == 2
        if two_d:
            t_seq = t_seq.transpose()
        n_t = len(t_seq)
        fmask = self._frame_mask(latent_mask, x_start)
        tab = self.tables(device)
        acc = th.zeros(n_t, B, 3, device=device, dtype=th.float64)
        x_t = th.empty_like(x_start)
        for i, t in enumerate(t_seq):
            t_batch = th.tensor(t, device=device).long() if two_d else th.full((B,), int(t), device=device, dtype=th.long)
            noise = th.randn_like(x_start)             # one draw per t (:970)
            with th.cuda.device(device), th.no_grad():
                ops.q_sample(x_start, noise, t_batch, tab, out=x_t)
                eps, _ = self._eps(model, x_t, t_batch, model_kwargs)
                ops.vb_terms(x_start, x_t, eps, noise, t_batch, tab, fmask, clip_denoised, acc[i])
        acc = acc.permute(1, 0, 2).float()
        prior = self._prior_bpd(x_start, latent_mask=latent_mask)
        vb = acc[:, :, 0].contiguous()
        return {'total_bpd': vb.sum(dim=1) + prior, 'prior_bpd': prior, 'vb': vb,
                'xstart_mse': acc[:, :, 1].contiguous(), 'mse': acc[:, :, 2].contiguous()}

    def calc_bpd_loop(self, model, x_start, clip_denoised=True, model_kwargs=None, latent_mask=None):
        return self.calc_bpd_loop_subsampled(model, x_start, clip_denoised=clip_denoised, model_kwargs=model_kwargs,
                                             latent_mask=latent_mask, t_seq=list(range(self.num_timesteps))[::-1])

    def training_losses(self, *args, **kwargs):
        raise NotImplementedError('training (backward pass) is outside the B200 inference hot path')
