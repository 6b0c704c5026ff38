"""Golden strings for the result-directory / run-identifier naming, produced by the reference's own
improved_diffusion/test_util.py (its unavailable imports stubbed).  Run in the build container only:
    python oracle/make_golden_formats.py   ->  tests/golden/formats.json
Test infrastructure: not imported by the product."""
import argparse
import json
import os
import sys
import tempfile
import types

import torch

sys.path.insert(0, '/root/reference')
for name in ('imageio', 'filelock', 'PIL', 'lpips', 'blobfile', 'mpi4py'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['filelock'].FileLock = object
sys.modules['PIL'].Image = types.ModuleType('Image')
sys.modules['PIL.Image'] = sys.modules['PIL'].Image
sys.modules['lpips'].LPIPS = type('LPIPS', (), {})
sys.modules['mpi4py'].MPI = None
from improved_diffusion import test_util as T  # noqa: E402

CASES_PATH = [
    dict(use_ddim=False, timestep_respacing='', eval_dir=None, checkpoint='ckpts-checkpoints/abcdefg/ema_0.9999_550000.pt', step=None, postfix=''),
    dict(use_ddim=True, timestep_respacing='100', eval_dir=None, checkpoint='my_checkpoint_root/run7/sub/ema_latest.pt', step=123000, postfix=''),
    dict(use_ddim=False, timestep_respacing='ddim25', eval_dir=None, checkpoint='x/checkpoints/a/b/model_latest.pt', step=7, postfix='_p'),
    dict(use_ddim=True, timestep_respacing='', eval_dir='some/eval/dir', checkpoint='x/checkpoints/a/model010.pt', step=None, postfix=''),
]
CASES_ID = [
    dict(inference_mode='autoreg', max_frames=20, step_size=7, T=500, obs_length=36),
    dict(inference_mode='exp-past', optimality='linspace-t', max_frames=20, step_size=8, T=300, obs_length=36,
         dataset_partition='train'),
    dict(inference_mode='independent', optimality=None, max_frames=10, step_size=5, T=30, obs_length=5,
         use_gradient_method=True, override_dataset='carla', dataset_partition='test'),
]


def main():
    out = dict(paths=[], ids=[])
    with tempfile.TemporaryDirectory() as tmp:
        for c in CASES_PATH:
            ck = os.path.join(tmp, c['checkpoint'])
            os.makedirs(os.path.dirname(ck), exist_ok=True)
            torch.save(dict(state_dict={}, config={}, step=c['step']), ck)
            args = argparse.Namespace(use_ddim=c['use_ddim'], timestep_respacing=c['timestep_respacing'],
                                      eval_dir=c['eval_dir'], checkpoint_path=ck)
            got = str(T.get_model_results_path(args, postfix=c['postfix']))
            out['paths'].append(dict(case=c, expect=got.replace(tmp, '<tmp>')))
    for c in CASES_ID:
        for postfix in ('', '_x'):
            out['ids'].append(dict(case=c, postfix=postfix,
                                   expect=T.get_eval_run_identifier(argparse.Namespace(**c), postfix=postfix)))
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'formats.json')
    json.dump(out, open(dst, 'w'), indent=1)
    print('wrote', dst, len(out['paths']), len(out['ids']))


if __name__ == '__main__':
    main()
