"""Factories with the reference's names and kwargs (`improved_diffusion/script_util.py`:
defaults :15-57, `create_video_model_and_diffusion` :110-181, `create_video_model` :229-300,
`create_gaussian_diffusion` :405-436, argparse helpers :439-467), building the B200 model."""
import argparse
import random

import numpy as np
import torch

from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps
from .unet import CondMargVideoModel, UNetVideoModel

NUM_CLASSES = 1000
CHANNEL_MULT = {256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4), 32: (1, 2, 2, 2)}


def model_and_diffusion_defaults():
    return dict(image_size=-1, num_channels=128, num_res_blocks=2, num_heads=4, num_heads_upsample=-1,
                attention_resolutions='16,8', dropout=0.0, learn_sigma=False, sigma_small=False, class_cond=False,
                diffusion_steps=1000, noise_schedule='linear', timestep_respacing='', use_kl=False,
                predict_xstart=False, rescale_timesteps=True, rescale_learned_sigmas=True, use_checkpoint=False,
                use_scale_shift_norm=True, use_spatial_encoding=False)


def video_model_and_diffusion_defaults():
    d = model_and_diffusion_defaults()
    d.update(T=-1, use_spatial_encoding=True, use_frame_encoding=False, cross_frame_attention=True, do_cond_marg=True,
             enforce_position_invariance=False, temporal_augment_type='add_manyhead_presoftmax_time', use_rpe_net=True,
             cond_emb_type='channel', rp_alpha=None, rp_beta=None, rp_gamma=None,
             allow_interactions_between_padding=True)
    return d


_MODEL_KEYS = ('learn_sigma', 'class_cond', 'use_checkpoint', 'attention_resolutions', 'num_heads',
               'num_heads_upsample', 'use_scale_shift_norm', 'dropout', 'use_spatial_encoding', 'use_frame_encoding',
               'cross_frame_attention', 'do_cond_marg', 'enforce_position_invariance', 'temporal_augment_type',
               'use_rpe_net', 'rp_alpha', 'rp_beta', 'rp_gamma', 'cond_emb_type', 'allow_interactions_between_padding')
_DIFFUSION_KEYS = ('learn_sigma', 'sigma_small', 'noise_schedule', 'use_kl', 'predict_xstart', 'rescale_timesteps',
                   'rescale_learned_sigmas', 'timestep_respacing')


def create_video_model_and_diffusion(T, image_size, num_channels, num_res_blocks, diffusion_steps, compute_dtype=None,
                                     **kw):
    model = create_video_model(T, image_size, num_channels, num_res_blocks, compute_dtype=compute_dtype,
                               **{k: kw[k] for k in _MODEL_KEYS})
    diffusion = create_gaussian_diffusion(steps=diffusion_steps, **{k: kw[k] for k in _DIFFUSION_KEYS})
    return model, diffusion


def create_video_model(T, image_size, num_channels, num_res_blocks, learn_sigma, class_cond, use_checkpoint,
                       attention_resolutions, num_heads, num_heads_upsample, use_scale_shift_norm, dropout,
                       use_spatial_encoding, use_frame_encoding, cross_frame_attention, do_cond_marg,
                       enforce_position_invariance, temporal_augment_type, use_rpe_net, rp_alpha, rp_beta, rp_gamma,
                       cond_emb_type, allow_interactions_between_padding, compute_dtype=None):
    if image_size not in CHANNEL_MULT:
        raise ValueError(f'unsupported image size: {image_size}')
    attention_ds = tuple(image_size // int(res) for res in attention_resolutions.split(','))
    bucket_params = dict(alpha=rp_alpha, beta=rp_beta, gamma=rp_gamma) if any([rp_alpha, rp_beta, rp_gamma]) else None
    cls = CondMargVideoModel if do_cond_marg else UNetVideoModel
    extra = dict(cond_emb_type=cond_emb_type) if do_cond_marg else {}
    return cls(T=T, in_channels=3, model_channels=num_channels, out_channels=(3 if not learn_sigma else 6),
               num_res_blocks=num_res_blocks, attention_resolutions=attention_ds, dropout=dropout,
               channel_mult=CHANNEL_MULT[image_size], num_classes=(NUM_CLASSES if class_cond else None),
               use_checkpoint=use_checkpoint, num_heads=num_heads, num_heads_upsample=num_heads_upsample,
               use_scale_shift_norm=use_scale_shift_norm, use_spatial_encoding=use_spatial_encoding,
               use_frame_encoding=use_frame_encoding, cross_frame_attention=cross_frame_attention,
               enforce_position_invariance=enforce_position_invariance, image_size=image_size,
               temporal_augment_type=temporal_augment_type, use_rpe_net=use_rpe_net, bucket_params=bucket_params,
               allow_interactions_between_padding=allow_interactions_between_padding,
               compute_dtype=compute_dtype, **extra)


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule='linear',
                              use_kl=False, predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing=''):
    betas = gd.get_named_beta_schedule(noise_schedule, steps)
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    elif rescale_learned_sigmas:
        loss_type = gd.LossType.RESCALED_MSE
    else:
        loss_type = gd.LossType.MSE
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing or [steps]), betas=betas,
                           model_mean_type=gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON,
                           model_var_type=var_type, loss_type=loss_type, rescale_timesteps=rescale_timesteps)


def add_dict_to_argparser(parser, default_dict):
    for k, v in default_dict.items():
        v_type = str if v is None else (str2bool if isinstance(v, bool) else type(v))
        parser.add_argument(f'--{k}', default=v, type=v_type)


def args_to_dict(args, keys):
    backups = {'allow_interactions_between_padding': True}
    return {k: getattr(args, k) if hasattr(args, k) else backups[k] for k in keys}


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('boolean value expected')


def set_random_seed(seed, deterministic=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
