"""ctypes binding of ``libhumanliff_b200.so`` (the C ABI declared in ``include/humanliff_b200.h``).

There is no fallback: if the shared library is missing or a symbol is absent the import of the
compute path fails loudly.  ``HL`` is the loaded library with argtypes set; ``check`` turns a
negative status into a ``RuntimeError`` carrying ``hl_last_error()``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HL_LIB") or os.path.join(_HERE, "libhumanliff_b200.so")   # $HL_LIB: experiment builds

c_int, c_i64, c_u64, c_f, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/humanliff_b200.h one to one
SIGNATURES = {
    "hl_version": (c_int, []),
    "hl_last_error": (ctypes.c_char_p, []),
    "hl_conv2d_uses_tensor_cores": (c_int, [c_int] * 11),
    "hl_conv_set_tuning": (c_int, [c_int] * 5),
    "hl_conv_set_tuning2": (c_int, [c_int] * 3),
    "hl_conv_set_workspace": (c_int, [c_p, c_i64, c_p]),
    "hl_conv_set_split": (c_int, [c_int]),
    "hl_conv_set_split_reduce": (c_int, [c_int]),
    "hl_conv2d_plan_info": (c_int, [c_int] * 10 + [c_i64, c_p]),
    "hl_conv_set_profile": (c_int, [c_p]),
    "hl_nchw_to_nhwc": (c_int, [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "hl_nhwc_to_nchw": (c_int, [c_p, c_int, c_p, c_int, c_int, c_int, c_p]),
    "hl_nhwc_to_nchw_sum2": (c_int, [c_p, c_int, c_int, c_p, c_int, c_int, c_int, c_p]),
    "hl_upsample2x": (c_int, [c_p, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "hl_cast_operand": (c_int, [c_p, c_int, c_p, c_int, c_int, c_int, c_i64, c_int, c_p]),
    "hl_zero": (c_int, [c_p, c_i64, c_p]),
    "hl_timestep_embedding": (c_int, [c_p, c_p, c_int, c_int, c_p, c_p]),
    "hl_linear_small": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p, c_p, c_p]),
    "hl_gn_stats": (c_int, [c_p, c_int, c_int, c_int, c_int, c_p, c_int, c_p]),
    "hl_gn_set_tuning": (c_int, [c_int]),
    "hl_gn_skip_supported": (c_int, [c_int] * 4),
    "hl_gn_skip_set_profile": (c_int, [c_p]),
    "hl_gn_skip": (c_int, [c_p, c_int, c_p, c_int, c_p, c_p, c_p, c_int, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int,
                           c_int, c_f, c_p]),
    "hl_gn_apply": (c_int, [c_p, c_int, c_p, c_int, c_p, c_p, c_p, c_int, c_p, c_int, c_int, c_p, c_int,
                            c_int, c_int, c_int, c_int, c_f, c_int, c_int, c_p]),
    "hl_conv_cout_pad": (c_int, [c_int]),
    "hl_conv2d": (c_int, [c_p, c_int, c_int, c_p, c_p, c_p, c_int, c_p, c_int, c_p, c_int, c_int, c_int,
                          c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "hl_conv2d_dual": (c_int, [c_p, c_int, c_int, c_p, c_p, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int, c_p, c_int,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "hl_set_pdl": (c_int, [c_int]),
    "hl_pdl_barrier": (None, []),
    "hl_launch_count": (c_i64, []),
    "hl_attention": (c_int, [c_p, c_int, c_int, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "hl_ddpm_step": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_i64, c_int, c_p]),
    "hl_ddim_step": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_i64, c_int, c_p]),
    "hl_randn": (c_int, [c_p, c_i64, c_p, c_u64, c_u64, c_i64, c_p]),
    "hl_ddpm_step_rng": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_p, c_p, c_int, c_i64, c_int, c_p, c_u64, c_u64, c_i64, c_p]),
    "hl_ddpm_posterior": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_p, c_p, c_int, c_i64, c_int, c_p, c_u64, c_u64, c_i64, c_p]),
    "hl_loop_advance": (c_int, [c_p, c_p, c_p, c_f, c_int, c_p, c_p]),
    "hl_triplane_to_texels": (c_int, [c_p, c_p, c_int, c_p]),
    "hl_render_rays": (c_int, [c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_u64, c_p, c_p, c_p, c_p,
                               c_i64, c_int, c_p]),
    "hl_smpl_vertex_tables": (c_int, [c_p, c_p, c_p, c_int, c_int, c_p, c_p, c_int, c_int, c_p, c_int, c_int, c_p, c_p, c_p]),
    "hl_render_rays_canon": (c_int, [c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_u64, c_p, c_p, c_p, c_int, c_int,
                                     c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_p]),
    "hl_render_rays_tc5_canon": (c_int, [c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_u64, c_p, c_p, c_p, c_int, c_int,
                                         c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_int, c_p]),
    "hl_canonical_points": (c_int, [c_p, c_p, c_i64, c_p, c_p, c_int, c_int, c_p, c_p, c_p, c_p, c_p]),
    "hl_density_grid_canon": (c_int, [c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_p, c_p, c_int, c_p, c_p]),
    "hl_render_set_profile": (c_int, [c_p]),
    "hl_density_grid_tc": (c_int, [c_p, c_int, c_p, c_p, c_p, c_int, c_p, c_p]),
    "hl_render_rays_tc5": (c_int, [c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_u64, c_p, c_int, c_p, c_p, c_p,
                                   c_i64, c_int, c_int, c_p]),
    "hl_density_grid_tc5": (c_int, [c_p, c_int, c_p, c_p, c_int, c_int, c_p, c_p]),
    "hl_render5_set_profile": (c_int, [c_p]),
    "hl_triplane_to_quads": (c_int, [c_p, c_p, c_int, c_p]),
    "hl_render_rays_tc": (c_int, [c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_u64, c_p, c_p, c_p, c_p,
                                  c_i64, c_int, c_p]),
}

DT_F32, DT_F16 = 0, 1

# offsets of the packed renderer MLP (HL_MLP_* in the header)
MLP_W0 = 0
MLP_B0 = MLP_W0 + 27 * 128
MLP_W1 = MLP_B0 + 128
MLP_B1 = MLP_W1 + 128 * 128
MLP_W2 = MLP_B1 + 128
MLP_B2 = MLP_W2 + 155 * 128
MLP_WA = MLP_B2 + 128
MLP_BA = MLP_WA + 128
MLP_WF = MLP_BA + 4
MLP_BF = MLP_WF + 128 * 128
MLP_WV = MLP_BF + 128
MLP_BV = MLP_WV + 155 * 64
MLP_WR = MLP_BV + 64
MLP_BR = MLP_WR + 64 * 4
MLP_PACK_FLOATS = MLP_BR + 4

# fp16 weight image of the tensor-core renderer (HL_MLP16_* in the header): (offset, rows, pitch)
MLP16_W0 = 0
MLP16_W1 = MLP16_W0 + 128 * 40
MLP16_W2 = MLP16_W1 + 128 * 136
MLP16_WF = MLP16_W2 + 128 * 168
MLP16_WV = MLP16_WF + 128 * 136
MLP16_HALVES = MLP16_WV + 64 * 136

SMPL_CLUSTERS = 128                                     # spatial clusters of the nearest-vertex search
smpl_consts = lambda J: 24 * J + 18 * (J - 1) + 28     # HL_SMPL_CONSTS(J)

MLP_TC5_BYTES = 16384 + 32768 + 16384 + 32768 + 32768 + 16384 + 16384 + 8192 + 392 * 4

CONV_FORCE_SIMT = 1
CONV_UPSAMPLE2X = 2
CONV_TF32 = 4
CONV_OUT_F16 = 8
CONV_SPLIT3 = 16
CONV_SPLIT2P = 32
CONV_OUT_F16_SPLIT = 64
CONV_SPLIT2A = 128
OP_TF32, OP_SCALED, OP_SPLIT, OP_RAW_SHIFT, OP_X_F16 = 1, 2, 4, 4, 8

_lib = None
launch_count = 0   # number of C-ABI compute calls issued (each is >= 1 kernel launch)


def load():
    """Load the shared library (once) and set every prototype.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m humanliff_b200.build` "
            "(there is no CPU / PyTorch fallback for the compute path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().hl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libhumanliff_b200 {what} failed ({status}): {msg}")


def call(name, *args):
    """Invoke one C-ABI entry point and raise on a non-zero status."""
    global launch_count
    launch_count += 1
    check(getattr(load(), name)(*args), name)
