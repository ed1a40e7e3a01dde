"""Factory functions with the reference's flag surface
(human_diffusion/improved_diffusion/script_util.py:11-39,42-150,260-298)."""
from . import diffusion as gd
from .diffusion import SpacedDiffusion, space_timesteps
from .unet import UNetModel

NUM_CLASSES = 4

_CHANNEL_MULT = {256: (1, 1, 2, 2, 4, 4), 224: (1, 1, 2, 2, 4, 4), 192: (1, 1, 2, 2, 4, 4),
                 128: (1, 1, 2, 2, 4, 4), 64: (1, 2, 3, 4), 32: (1, 2, 2, 2)}


def model_and_diffusion_defaults():
    """Same keys / values as script_util.py:11-39."""
    return dict(image_size=64, in_channels=3, num_channels=128, out_channels=3, num_res_blocks=2,
                num_heads=4, num_heads_upsample=-1, attention_resolutions="16,8", dropout=0.0,
                learn_sigma=False, sigma_small=False, class_cond=False, diffusion_steps=1000,
                noise_schedule="linear", timestep_respacing="", use_kl=False, predict_xstart=False,
                rescale_timesteps=True, rescale_learned_sigmas=True, use_checkpoint=False,
                use_scale_shift_norm=True, cond_type="controlnet", use_3d_aware=False)


def create_model(image_size, in_channels, num_channels, out_channels, num_res_blocks, learn_sigma,
                 class_cond, use_checkpoint, attention_resolutions, num_heads, num_heads_upsample,
                 use_scale_shift_norm, cond_type, use_3d_aware, dropout, precision="fp16"):
    if image_size not in _CHANNEL_MULT:
        raise ValueError(f"unsupported image size: {image_size}")
    attention_ds = tuple(image_size // int(res) for res in attention_resolutions.split(","))
    return UNetModel(in_channels=in_channels, model_channels=num_channels,
                     out_channels=(out_channels if not learn_sigma else out_channels * 2),
                     num_res_blocks=num_res_blocks, attention_resolutions=attention_ds, dropout=dropout,
                     channel_mult=_CHANNEL_MULT[image_size],
                     num_classes=(NUM_CLASSES if class_cond else None), use_checkpoint=use_checkpoint,
                     num_heads=num_heads, num_heads_upsample=num_heads_upsample,
                     use_scale_shift_norm=use_scale_shift_norm, cond_type=cond_type,
                     use_3d_aware=use_3d_aware, precision=precision)


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear",
                              use_kl=False, predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing=""):
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
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, timestep_respacing or [steps]), betas=betas,
        model_mean_type=(gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON),
        model_var_type=var_type, loss_type=loss_type, rescale_timesteps=rescale_timesteps)


def create_model_and_diffusion(image_size, class_cond, learn_sigma, sigma_small, in_channels, num_channels,
                               out_channels, num_res_blocks, num_heads, num_heads_upsample,
                               attention_resolutions, dropout, diffusion_steps, noise_schedule,
                               timestep_respacing, use_kl, predict_xstart, rescale_timesteps,
                               rescale_learned_sigmas, use_checkpoint, use_scale_shift_norm, cond_type,
                               use_3d_aware, precision="fp16"):
    """The 23 keyword flags of script_util.py:42-66 (+ ``precision``); returns (model, diffusion)."""
    model = create_model(image_size, in_channels, num_channels, out_channels, num_res_blocks,
                         learn_sigma=learn_sigma, class_cond=class_cond, use_checkpoint=use_checkpoint,
                         attention_resolutions=attention_resolutions, num_heads=num_heads,
                         num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
                         cond_type=cond_type, use_3d_aware=use_3d_aware, dropout=dropout,
                         precision=precision)
    diffusion = create_gaussian_diffusion(steps=diffusion_steps, learn_sigma=learn_sigma,
                                          sigma_small=sigma_small, noise_schedule=noise_schedule,
                                          use_kl=use_kl, predict_xstart=predict_xstart,
                                          rescale_timesteps=rescale_timesteps,
                                          rescale_learned_sigmas=rescale_learned_sigmas,
                                          timestep_respacing=timestep_respacing)
    return model, diffusion


def production_flags(timestep_respacing=""):
    """Flags of human_diffusion/triplane_scripts/SynBody_triplane_sample_layered_*.sh:10,24-26."""
    kw = model_and_diffusion_defaults()
    kw.update(image_size=256, in_channels=27, num_channels=192, out_channels=27, num_res_blocks=3,
              num_heads=4, attention_resolutions="32,16,8", class_cond=True, learn_sigma=False,
              noise_schedule="linear", diffusion_steps=1000, timestep_respacing=timestep_respacing,
              rescale_timesteps=False, rescale_learned_sigmas=False, use_scale_shift_norm=True,
              cond_type="controlnet", use_3d_aware=False)
    return kw
