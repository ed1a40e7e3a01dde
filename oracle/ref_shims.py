"""Import the UNMODIFIED reference from /root/reference on this container's CPU.
TEST INFRASTRUCTURE ONLY; used by make_goldens.py and the (optional) reference-vs-oracle tests.
Nothing here runs on the GPU box -- /root/reference does not exist there."""
import os
import sys
import types

REF_ROOT = os.environ.get("HUMANLIFF_REF", "/root/reference")


def use(root):
    """Point the shims at another copy of the unmodified reference (``oracle/_ref``, staged by
    ``oracle/build_ref.py`` for the GPU box where /root/reference does not exist)."""
    global REF_ROOT
    REF_ROOT = root


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "human_diffusion", "improved_diffusion"))


def import_diffusion():
    """-> (script_util module) of the reference's improved_diffusion package."""
    p = os.path.join(REF_ROOT, "human_diffusion")
    if p not in sys.path:
        sys.path.insert(0, p)
    from improved_diffusion import script_util   # noqa
    return script_util


def import_hd_renderer():
    """-> the reference's human_diffusion/NeRF/renderer.py module, with stub modules for the absent
    mcubes / pytorch3d (never touched when use_canonical_space=False) and anomaly mode switched back off."""
    import torch
    for name in ("mcubes", "pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name.endswith("knn"):
                m.knn_points = None
            sys.modules[name] = m
    p = os.path.join(REF_ROOT, "human_diffusion")
    if p not in sys.path:
        sys.path.insert(0, p)
    from NeRF import renderer as hd_renderer    # noqa
    torch.autograd.set_detect_anomaly(False)      # fields.py:2 turns it on at import
    return hd_renderer


def import_rn_renderer():
    """-> the reference's recon_NeRF/lib/renderer.py module.  Besides the mcubes / pytorch3d stubs it needs the
    SMPL asset load of Renderer.__init__ (renderer.py:45-48: read_pickle + SMPL_to_tensor on
    torch.cuda.current_device()) neutralised -- none of it is touched when use_canonical_space=False."""
    import torch
    for name in ("mcubes", "pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name.endswith("knn"):
                m.knn_points = None
            sys.modules[name] = m
    p = os.path.join(REF_ROOT, "recon_NeRF")
    if p not in sys.path:
        sys.path.insert(0, p)
    from lib import renderer as rn_renderer     # noqa
    torch.autograd.set_detect_anomaly(False)
    rn_renderer.read_pickle = lambda path: {}
    rn_renderer.SMPL_to_tensor = lambda params, device=None: {"f": None}
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0
    return rn_renderer
