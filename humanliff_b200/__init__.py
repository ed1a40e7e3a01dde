"""humanliff_b200 -- B200-native (sm_100a) implementation of HumanLiff's two hot paths:
the improved_diffusion UNet denoising loop and the tri-plane volume renderer.

Public surface mirrors the reference:
    create_model_and_diffusion, model_and_diffusion_defaults, create_gaussian_diffusion   (script_util.py)
    UNetModel                                                                             (unet.py)
    GaussianDiffusion, SpacedDiffusion, space_timesteps                                   (gaussian_diffusion.py, respace.py)
    Renderer, render                                                                      (renderer.py, run_nerf_batch.py)
    all_gather_samples                                                                    (triplane_sample_layered.py:211-219)
    sample_layer, sample_all_layers                                                       (triplane_sample_layered.py:110-151,229-244)
"""
from .diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType, SpacedDiffusion,
                        get_named_beta_schedule, space_timesteps)
from .factory import (create_gaussian_diffusion, create_model, create_model_and_diffusion,
                      model_and_diffusion_defaults, production_flags)
from .unet import UNetModel
from .renderer import Renderer, render
from .dist import all_gather_samples, shard_batch
from .layered import sample_all_layers, sample_layer

__all__ = ["GaussianDiffusion", "SpacedDiffusion", "space_timesteps", "get_named_beta_schedule",
           "ModelMeanType", "ModelVarType", "LossType", "create_model_and_diffusion", "create_model",
           "create_gaussian_diffusion", "model_and_diffusion_defaults", "production_flags", "UNetModel",
           "Renderer", "render", "all_gather_samples", "shard_batch", "sample_layer", "sample_all_layers"]
