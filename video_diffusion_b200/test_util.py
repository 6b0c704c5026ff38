"""File formats and naming at the edges of the hot path, as the reference's evaluation scripts expect them
(improved_diffusion/test_util.py:31-132, scripts/video_nll.py:94-138): checkpoint loader, result-directory
naming, run identifier, per-video ELBO pickles.  Host-side only; nothing here touches the GPU kernels."""
import argparse
import os
import pickle
from pathlib import Path

import numpy as np
import torch

from .script_util import args_to_dict, create_video_model_and_diffusion, video_model_and_diffusion_defaults


def _load(checkpoint_path, trust_pickle=False):
    """A checkpoint holds tensors, a plain config dict and an int: `weights_only=True` loads that without executing
    pickled code.  Checkpoints whose config was saved as an argparse.Namespace (or anything else the safe loader
    rejects) need the explicit `trust_pickle=True` opt-in."""
    try:
        return torch.load(checkpoint_path, map_location='cpu', weights_only=True)
    except pickle.UnpicklingError:
        if not trust_pickle:
            raise
        return torch.load(checkpoint_path, map_location='cpu', weights_only=False)


def load_checkpoint(checkpoint_path, device, use_ddim=False, timestep_respacing='', compute_dtype=None,
                    trust_pickle=False):
    """Reads a reference checkpoint `{'state_dict', 'config', 'step'}` and builds the B200 model + diffusion from its
    config (test_util.py:31-62).  Configs written before `enforce_position_invariance` / `cond_emb_type` existed get
    the reference's back-compat defaults.  Returns ((model, diffusion), model_args)."""
    default_model_configs = {'enforce_position_invariance': False, 'cond_emb_type': 'channel'}
    data = _load(checkpoint_path, trust_pickle)
    state_dict = data['state_dict']
    model_args = dict(vars(data['config']) if isinstance(data['config'], argparse.Namespace) else data['config'])
    model_args.update({'use_ddim': use_ddim, 'timestep_respacing': timestep_respacing})
    for k, v in default_model_configs.items():
        model_args.setdefault(k, v)
    model_args = argparse.Namespace(**model_args)
    kwargs = args_to_dict(model_args, video_model_and_diffusion_defaults().keys())
    if compute_dtype is not None:
        kwargs['compute_dtype'] = compute_dtype
    model, diffusion = create_video_model_and_diffusion(**kwargs)
    model.load_state_dict(state_dict)
    model = model.to(device)
    model.eval()
    return (model, diffusion), model_args


def get_model_results_path(args, postfix=''):
    """`results/<checkpoint subpath after the first *checkpoint* directory>/<checkpoint stem>[_<step>][_ddim][_respaceN]`
    (test_util.py:65-107); `args.eval_dir`, when set, wins."""
    if args.use_ddim:
        postfix += '_ddim'
    if args.timestep_respacing != '':
        postfix += '_' + f'respace{args.timestep_respacing}'
    if args.eval_dir is not None:
        return Path(args.eval_dir)
    checkpoint_path = Path(args.checkpoint_path)
    name = f'{checkpoint_path.stem}'
    if name.endswith('latest'):
        step = _load(args.checkpoint_path, getattr(args, 'trust_pickle', False))['step']
        name += f'_{step}'
    if postfix != '':
        name += postfix
    path = None
    for idx, part in enumerate(checkpoint_path.parts):
        if 'checkpoint' in part:
            path = Path(*(checkpoint_path.parts[idx + 1:]))
            break
    assert path is not None, 'the checkpoint path must contain a directory whose name includes "checkpoint"'
    return Path('results') / path.parent / name


def get_eval_run_identifier(args, postfix=''):
    """`<mode>[_optimal-X]_<max_frames>_<step_size>_<T>_<obs_length>` with the reference's prefixes (test_util.py:110-132)."""
    res = args.inference_mode
    if getattr(args, 'optimality', None) is not None:
        res += f'_optimal-{args.optimality}'
    res += f'_{args.max_frames}_{args.step_size}_{args.T}_{args.obs_length}'
    if getattr(args, 'dataset_partition', None) == 'train':
        res = 'trainset_' + res
    if getattr(args, 'use_gradient_method', False):
        res = 'gradientmethod_' + res
    if getattr(args, 'override_dataset', None) is not None:
        res = f'{args.override_dataset}_' + res
    if postfix != '':
        res += postfix
    return res


def save_elbos(eval_dir, returns, dataset_indices, postfix=''):
    """`returns`: one `run_bpd_evaluation` result per index type (each a dict of arrays with leading batch dim).  Writes
    `elbos/elbo_<dataset idx><postfix>.pkl` per video holding `{key: array stacked over index types}` exactly like
    scripts/video_nll.py:126-137, and returns the paths."""
    out_dir = Path(eval_dir) / 'elbos'
    os.makedirs(out_dir, exist_ok=True)
    stacked = {k: np.stack([np.asarray(r[k]) for r in returns], axis=1) for k in returns[0].keys()}
    paths = []
    for j, idx in enumerate(dataset_indices):
        fname = out_dir / f'elbo_{idx}{postfix}.pkl'
        with open(fname, 'wb') as f:
            pickle.dump({k: v[j] for k, v in stacked.items()}, f)
        paths.append(fname)
    return paths
