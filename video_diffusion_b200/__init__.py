"""B200-native hot path of cliangyu/video-diffusion: the video U-Net denoising forward and the
diffusion sampling / ELBO loops, as hand-written sm_100a CUDA behind the reference's Python API."""
from . import gaussian_diffusion, respace, script_util, unet  # noqa: F401
from .script_util import (create_gaussian_diffusion, create_video_model,  # noqa: F401
                          create_video_model_and_diffusion, video_model_and_diffusion_defaults)

__all__ = ['gaussian_diffusion', 'respace', 'script_util', 'unet', 'create_gaussian_diffusion', 'create_video_model',
           'create_video_model_and_diffusion', 'video_model_and_diffusion_defaults']
