"""Layer-wise generation: the outer loop of human_diffusion/scripts/triplane_sample_layered.py.

HumanLiff generates a clothed human layer by layer (body -> pants -> shirt -> shoes): layer k is sampled
with class label ``y = k`` and the finished tri-plane of layer k-1 as the ControlNet condition ``x_cond``
(zeros for layer 0) -- ``triplane_sample_layered.py:110-151``.  The reference runs one process per layer and
hands the tri-planes over through ``.npz`` files (``arr_0`` = samples ``[N,27,H,W]`` fp32, ``arr_1`` = labels,
``:229-244``; re-read at ``:131-132``), reloading the 2 GB model each time.  Here:

* :func:`sample_layer` is one such invocation (same hand-off file format, same naming), and
* :func:`sample_all_layers` keeps the four layers in ONE process: the device tensor of layer k-1 is the
  ``x_cond`` of layer k directly -- no disk round trip, no model reload.  The ``.npz`` files hold fp32, so the
  two routes give bit-identical tri-planes (tests/test_unet_gpu.py::test_layered_handoff).

SURVEY.md 8(f) rank 4 / 8(d) config 4.
"""
import os

import numpy as np
import torch

from .dist import all_gather_samples

# file-name stems of triplane_sample_layered.py:232-239
LAYER_NAMES = ("person", "person_pant", "person_pant_shirt", "person_pant_shirt_shoes")
NUM_LAYERS = len(LAYER_NAMES)


def layer_npz_path(out_dir, layer_index, shape, suffix, start_id=0):
    """``samples_<stem>_<NxCxHxW>_<suffix>_start_id_<id>.npz`` (triplane_sample_layered.py:229-239)."""
    shape_str = "x".join(str(int(v)) for v in shape)
    return os.path.join(out_dir, f"samples_{LAYER_NAMES[layer_index]}_{shape_str}_{suffix}_start_id_{start_id}.npz")


def save_layer_npz(path, samples, labels=None):
    """np.savez(out_path, arr, label_arr) -> keys arr_0, arr_1 (triplane_sample_layered.py:241-244)."""
    arr = samples.detach().cpu().numpy() if torch.is_tensor(samples) else np.asarray(samples)
    if labels is None:
        np.savez(path, arr)
    else:
        lab = labels.detach().cpu().numpy() if torch.is_tensor(labels) else np.asarray(labels)
        np.savez(path, arr, lab)
    return path


def load_layer_cond(path, start, count, device):
    """``th.from_numpy(np.load(npz).f.arr_0.astype(float32))[start:start+count]`` (triplane_sample_layered.py:131-132)."""
    with np.load(path) as z:
        arr = z["arr_0"].astype(np.float32)
    if start + count > arr.shape[0]:
        raise ValueError(f"{path}: holds {arr.shape[0]} samples, rows [{start}, {start + count}) requested")
    return torch.from_numpy(arr[start:start + count]).to(device)


def sample_layer(model, diffusion, layer_index, batch_size, x_cond=None, sample_npz=None, count=0, image_size=256,
                 shape=None, use_ddim=False, clip_denoised=True, class_cond=True, noise=None, step_noise=None,
                 device=None):
    """One pass of the ``while`` body at triplane_sample_layered.py:110-151 for one batch.

    ``x_cond``: tensor (kept on the device from the previous layer), or ``sample_npz`` = the previous layer's
    hand-off file (rows ``[count, count+batch_size)`` are used, as the reference does), or neither for layer 0
    (zeros).  Returns ``(sample [B,C,H,W], classes [B] int64)``."""
    if not 0 <= layer_index < NUM_LAYERS:
        raise ValueError(f"layer_index must be in [0, {NUM_LAYERS}), got {layer_index}")
    if device is None:
        device = next(model.parameters()).device
    if shape is None:
        shape = (batch_size, model.out_channels, image_size, image_size)            # :134-151
    shape = tuple(int(v) for v in shape)
    if shape[0] != batch_size:
        raise ValueError("shape[0] must equal batch_size")
    model_kwargs = {}
    classes = torch.full((batch_size,), layer_index, dtype=torch.int64, device=device)    # :113-117
    if class_cond:
        model_kwargs["y"] = classes
    if x_cond is None:
        if layer_index == 0 or sample_npz is None:
            if layer_index != 0:
                raise ValueError("layers 1..3 are conditioned on the previous layer: pass x_cond or sample_npz")
            x_cond = torch.zeros(shape, device=device)                                    # :124-129
        else:
            x_cond = load_layer_cond(sample_npz, count, batch_size, device)              # :130-132
    if tuple(x_cond.shape) != shape:
        raise ValueError(f"x_cond shape {tuple(x_cond.shape)} != sample shape {shape}")
    fn = diffusion.ddim_sample_loop if use_ddim else diffusion.p_sample_loop              # :118-120
    kw = dict(x_cond=x_cond, clip_denoised=clip_denoised, model_kwargs=model_kwargs, noise=noise)
    if step_noise is not None:
        kw["step_noise"] = step_noise
    sample = fn(model, shape, **kw)
    return sample, classes


def sample_all_layers(model, diffusion, batch_size, num_layers=NUM_LAYERS, image_size=256, shape=None, use_ddim=False,
                      clip_denoised=True, noise=None, step_noise=None, out_dir=None, suffix="b200", start_id=0,
                      gather=False, device=None):
    """All clothing layers of ``batch_size`` humans in one process: ``x_cond_0 = 0``, ``x_cond_k = sample_{k-1}``,
    ``y = k``.

    ``noise`` / ``step_noise``: optional callables ``k -> x_T`` and ``k -> (i -> z_i)`` that inject the Gaussians of
    layer k (parity tests).  ``out_dir``: also write the reference's hand-off files (rank 0 when ``gather``).
    ``gather``: all-gather each finished layer over the default process group exactly once per layer
    (``triplane_sample_layered.py:211-219``) and return the gathered tensors.  Returns a list of
    ``(samples, labels)`` per layer."""
    outs = []
    x_cond = None
    for k in range(num_layers):
        sample, classes = sample_layer(
            model, diffusion, k, batch_size, x_cond=x_cond, image_size=image_size, shape=shape, use_ddim=use_ddim,
            clip_denoised=clip_denoised, noise=noise(k) if noise is not None else None,
            step_noise=step_noise(k) if step_noise is not None else None, device=device)
        x_cond = sample                                         # stays in HBM: the next layer's condition
        g_s, g_c = all_gather_samples(sample, classes) if gather else (sample, classes)
        if out_dir is not None and (not gather or not torch.distributed.is_initialized()
                                    or torch.distributed.get_rank() == 0):
            os.makedirs(out_dir, exist_ok=True)
            save_layer_npz(layer_npz_path(out_dir, k, g_s.shape, suffix, start_id), g_s, g_c)
        outs.append((g_s, g_c))
    return outs
