"""CPU oracle for the HumanLiff hot paths.  TEST INFRASTRUCTURE ONLY.

Nothing under ``humanliff_b200/`` may import this package.  The only permitted
users are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker or as
the reported CPU baseline, never as the product path.

What it is: an independent functional restatement (torch CPU fp32 tensors, numpy
float64 schedule tables) of the reference algorithms

* ``unet_oracle``       -- ``UNetModel.forward``         (human_diffusion/improved_diffusion/unet.py:550-615)
* ``diffusion_oracle``  -- ``SpacedDiffusion`` / ``p_sample`` / ``p_sample_loop``
                           (gaussian_diffusion.py:118-169,232-326,356-482; respace.py:7-122)
* ``render_oracle``     -- ``render()`` + ``Renderer.render`` (recon_NeRF/run_nerf_batch.py:29-67,
                           recon_NeRF/lib/renderer.py:142-295,504-581; human_diffusion/NeRF/renderer.py:234-281)

The arithmetic of the reference lives in PyTorch ATen (third-party, pinned
``pytorch==1.11.0`` in the reference README:39; this image has 2.11.0), so the
restatement uses the same ATen primitives (conv2d, group_norm, grid_sample ...)
as its arithmetic library, but re-derives the model structure from the state
dict alone and shares no code with the product.

Parity pin: the reference ships no tests / golden vectors.
The pin is therefore "reference source executed here under torch 2.11 CPU fp32":
``oracle/make_goldens.py`` imports the unmodified reference from
``/root/reference`` (with the import shims in ``oracle/ref_shims.py``), runs it on
seeded inputs and freezes the outputs under ``tests/golden/``; ``tests/test_oracle_*``
check this restatement against those frozen outputs.
"""
