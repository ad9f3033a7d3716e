"""ctypes binding of include/marf_b200.h.

The product path has NO fallback: if the shared library is missing, loading raises; if there is
no B200, `marf_create` fails.  Nothing here imports the oracle.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libmarf_b200.so")

MARF_ABI_VERSION = 2
MARF_MAX_LAYERS = 12
MASK_NONE, MASK_DISK, MASK_IMPLICIT = 0, 1, 2
FP32, BF16 = 0, 1
S_RGB, N_RGB, S_MASK, N_MASK, S_EDGE, N_EDGE, NONFINITE, BAD_INDEX, N_SUMS = 0, 1, 2, 3, 4, 5, 6, 7, 8

EXPORTS = [
    "marf_abi_version", "marf_create", "marf_destroy", "marf_last_error", "marf_step", "marf_step_forward",
    "marf_step_backward", "marf_render", "marf_sl3_to_SL3", "marf_warp_corners", "marf_warp_points", "marf_compute_edges",
    "marf_launch_count", "marf_workspace_bytes", "marf_tc_selftest", "marf_adam_step", "marf_debug_read_bf16", "marf_profile", "marf_profile_read", "marf_loss_scalars", "marf_peer_allreduce",
    "marf_forward_points", "marf_tf32_gemm",
]

_i32, _u32, _i64, _f32, _f64, _vp = C.c_int32, C.c_uint32, C.c_int64, C.c_float, C.c_double, C.c_void_p


class MarfConfig(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("device", _i32), ("precision", _i32),
        ("H", _i32), ("W", _i32), ("patch_H", _i32), ("patch_W", _i32), ("use_cropped", _i32),
        ("batch_global", _i32), ("batch", _i32), ("patch_offset", _i32), ("rows", _i32), ("row_offset", _i32),
        ("L", _i32), ("n_layers", _i32), ("layer_out", _i32 * MARF_MAX_LAYERS), ("skip_mask", _u32),
        ("c2f_enabled", _i32), ("c2f_start", _f32), ("c2f_end", _f32),
        ("mask_mode", _i32), ("mask_n_layers", _i32), ("mask_layer_out", _i32 * MARF_MAX_LAYERS),
        ("mask_uv_freqs", _i32), ("mask_embed_dim", _i32),
        ("use_edges", _i32), ("edge_label_channels", _i32),
        ("max_chunk_pixels", _i64),
        ("mask_n_vocab", _i32), ("reserved0", _i32),
    ]


class MarfStepIO(C.Structure):
    _fields_ = [
        ("mlp_w", C.POINTER(_vp)), ("mlp_b", C.POINTER(_vp)), ("warp", _vp),
        ("mask_w", C.POINTER(_vp)), ("mask_b", C.POINTER(_vp)), ("embed", _vp),
        ("rgb", _vp), ("masks", _vp), ("masks_eroded", _vp), ("edges", _vp), ("data_version", _i64),
        ("progress", _f32), ("c_rgb", _f32), ("c_mask", _f32), ("c_edge", _f32),
        ("norm_rgb", _f64), ("norm_edge", _f64),
        ("g_mlp_w", C.POINTER(_vp)), ("g_mlp_b", C.POINTER(_vp)), ("g_warp", _vp),
        ("g_mask_w", C.POINTER(_vp)), ("g_mask_b", C.POINTER(_vp)),
        ("rgb_pred", _vp), ("mask_pred", _vp), ("edge_pred", _vp), ("loss_sums", _vp),
    ]


class MarfRenderIO(C.Structure):
    _fields_ = [
        ("mlp_w", C.POINTER(_vp)), ("mlp_b", C.POINTER(_vp)), ("warp", _vp),
        ("n_patches", _i32), ("crop", _i32), ("progress", _f32), ("rgb", _vp),
    ]


class MarfAdamIO(C.Structure):
    _fields_ = [
        ("n_tensors", _i32), ("params", C.POINTER(_vp)), ("grads", C.POINTER(_vp)), ("exp_avg", C.POINTER(_vp)),
        ("exp_avg_sq", C.POINTER(_vp)), ("numel", C.POINTER(_i64)), ("lr", C.POINTER(_f32)),
        ("beta1", _f32), ("beta2", _f32), ("eps", _f32), ("step", _i64), ("zero_tensor", _i32), ("zero_count", _i32),
    ]


_lib = None


def load():
    """Load libmarf_b200.so (built in-tree by `python -m marf_b200.build`).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m marf_b200.build` (nvcc, sm_100a). "
            "marf_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.marf_abi_version.restype = C.c_int
    lib.marf_create.argtypes = [C.POINTER(MarfConfig), C.POINTER(_vp)]
    lib.marf_create.restype = C.c_int
    lib.marf_destroy.argtypes = [_vp]
    lib.marf_destroy.restype = C.c_int
    lib.marf_last_error.argtypes = [_vp]
    lib.marf_last_error.restype = C.c_char_p
    for name in ("marf_step", "marf_step_forward", "marf_step_backward"):
        fn = getattr(lib, name)
        fn.argtypes = [_vp, C.POINTER(MarfStepIO), _vp]
        fn.restype = C.c_int
    lib.marf_render.argtypes = [_vp, C.POINTER(MarfRenderIO), _vp]
    lib.marf_render.restype = C.c_int
    lib.marf_forward_points.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _f32, _vp, _vp]
    lib.marf_forward_points.restype = C.c_int
    lib.marf_sl3_to_SL3.argtypes = [_vp, _vp, _i32, _vp, _vp]
    lib.marf_sl3_to_SL3.restype = C.c_int
    lib.marf_warp_corners.argtypes = [_vp, _vp, _i32, _vp, _vp]
    lib.marf_warp_corners.restype = C.c_int
    lib.marf_warp_points.argtypes = [_vp, _vp, _vp, _i32, _i32, _vp, _vp]
    lib.marf_warp_points.restype = C.c_int
    lib.marf_compute_edges.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]
    lib.marf_compute_edges.restype = C.c_int
    lib.marf_tc_selftest.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]
    lib.marf_tc_selftest.restype = C.c_int
    lib.marf_tf32_gemm.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp,
                                   C.c_int, _vp]
    lib.marf_tf32_gemm.restype = C.c_int
    lib.marf_debug_read_bf16.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_longlong, _vp]
    lib.marf_debug_read_bf16.restype = C.c_int
    lib.marf_profile.argtypes = [_vp, C.c_int]
    lib.marf_profile.restype = C.c_int
    lib.marf_profile_read.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]
    lib.marf_profile_read.restype = C.c_int
    lib.marf_loss_scalars.argtypes = [_vp, _vp, C.c_double, C.POINTER(C.c_double), _vp, _vp]
    lib.marf_loss_scalars.restype = C.c_int
    lib.marf_peer_allreduce.argtypes = [C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.c_int, C.c_int, _vp, C.c_longlong, _u32, _vp]
    lib.marf_peer_allreduce.restype = C.c_int
    lib.marf_adam_step.argtypes = [_vp, C.POINTER(MarfAdamIO), _vp]
    lib.marf_adam_step.restype = C.c_int
    lib.marf_launch_count.argtypes = [_vp]
    lib.marf_launch_count.restype = _i64
    lib.marf_workspace_bytes.argtypes = [_vp]
    lib.marf_workspace_bytes.restype = _i64
    if lib.marf_abi_version() != MARF_ABI_VERSION:
        raise RuntimeError("libmarf_b200.so ABI version mismatch; rebuild with `python -m marf_b200.build -f`")
    _lib = lib
    return lib


class MarfError(RuntimeError):
    pass


def check(lib, handle, rc, what):
    if rc != 0:
        msg = lib.marf_last_error(handle)
        raise MarfError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr_array(tensors):
    """host array of device pointers for a list of tensors (or None)."""
    if tensors is None:
        return None
    arr = (_vp * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr
