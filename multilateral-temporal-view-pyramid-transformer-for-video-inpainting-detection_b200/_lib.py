"""ctypes binding of libmumpy_b200.so (the C ABI declared in include/mumpy_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a tensor is not on a CUDA device the
call fails loudly.  Build the library with `python __graft_entry__.py` (or `python -m mumpy_b200.build`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MUMPY_LIB: another build of the same ABI (development A/B runs on one box); the product always loads the in-tree library
LIB_PATH = os.environ.get("MUMPY_LIB") or os.path.join(_HERE, "libmumpy_b200.so")

_lib = None

vp, ci, cl, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float

# name -> argtypes, exactly the prototypes of include/mumpy_b200.h
SIGNATURES = {
    "mumpy_abi_version": [],
    "mumpy_init": [ci],
    "mumpy_set_pdl": [ci],
    "mumpy_set_f16_overflow_flag": [vp],
    "mumpy_set_gemm_pair_mode": [ci],
    "mumpy_set_attention_tc": [ci],
    "mumpy_set_gemm_tile": [ci],
    "mumpy_linear": [vp, cl, vp, vp, vp, vp, cl, cl, ci, ci, ci, ci, ci, vp],
    "mumpy_linear_dual": [vp, cl, vp, vp, vp, vp, vp, cl, cl, ci, ci, ci, ci, vp],
    "mumpy_layernorm": [vp, vp, vp, vp, ci, cl, ci, cf, vp],
    "mumpy_mlp_fused_supported": [ci],
    "mumpy_set_mlp_fused_shape": [ci],
    "mumpy_mlp_fused": [vp, vp, vp, cf, vp, vp, vp, vp, vp, cl, ci, ci, vp],
    "mumpy_ln_linear_supported": [ci, ci],
    "mumpy_set_ln_linear_pair_mode": [ci],
    "mumpy_ln_linear": [vp, vp, vp, cf, vp, vp, vp, cl, cl, ci, ci, ci, ci, vp],
    "mumpy_patch_merge_norm": [vp, vp, vp, vp, ci, ci, ci, ci, ci, cf, vp],
    "mumpy_window_attention": [vp, vp, vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_mha_short": [vp, vp, ci, cl, ci, ci, ci, vp],
    "mumpy_mha_short_probs": [vp, vp, ci, cl, ci, ci, ci, vp],
    "mumpy_tokenize": [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, cf, vp],
    "mumpy_patchify16": [vp, vp, ci, ci, ci, ci, ci, vp],
    "mumpy_faf_workspace_floats": [ci, ci],
    "mumpy_faf16_workspace_bytes": [ci, ci],
    "mumpy_groupnorm_workspace_floats": [ci, ci, ci, ci],
    "mumpy_faf": [vp, vp, vp, vp, ci, ci, ci, ci, ctypes.POINTER(ci), vp],
    "mumpy_faf16": [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ctypes.POINTER(ci), ci, vp],
    "mumpy_cva_offsets": [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_cva_sample": [vp, ci, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_cva_attention": [vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_cva_attention_probs": [vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_cva_residual": [vp, vp, vp, ci, ci, ci, ci, ci, vp],
    "mumpy_gather_rows": [vp, ci, vp, ci, cl, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_conv2d_nhwc": [vp, cl, vp, vp, vp, cl, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_conv2d_nhwc_bf16": [vp, cl, vp, vp, vp, vp, cl, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, cl, vp],
    "mumpy_conv2d_nhwc_cout1": [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_im2col_nhwc": [vp, cl, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_groupnorm_nhwc": [vp, vp, vp, vp, vp, cl, ci, ci, ci, ci, ci, cf, ci, ci, vp],
    "mumpy_resample_nhwc": [vp, vp, vp, vp, ci, cl, ci, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_mul_add": [vp, vp, vp, vp, ci, cl, vp],
    "mumpy_add": [vp, vp, vp, cl, vp],
    "mumpy_nchw_to_nhwc": [vp, vp, cl, ci, ci, ci, ci, ci, ci, vp],
    "mumpy_nhwc_to_nchw": [vp, cl, vp, ci, ci, ci, ci, vp],
    "mumpy_channel_group_mean": [vp, vp, cl, ci, ci, vp],
    "mumpy_assemble_clips": [vp, vp, vp, ci, ci, ci, ci, ctypes.POINTER(cf), ctypes.POINTER(cf), vp],
    "mumpy_resize_taps": [ci, ci, ci, vp, vp, ci, vp],
    "mumpy_resize_u8": [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp, ci, vp],
    "mumpy_mask_counts": [vp, vp, vp, vp, ci, ci, vp],
    "mumpy_cast16": [vp, vp, ci, cl, vp],
}


class MumpyError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MumpyError(
            "libmumpy_b200.so not found at %s -- build it first (python __graft_entry__.py); "
            "this package has no CPU/PyTorch fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.argtypes = argtypes
        fn.restype = ci
    for name in ("mumpy_faf_workspace_floats", "mumpy_faf16_workspace_bytes", "mumpy_groupnorm_workspace_floats"):
        getattr(lib, name).restype = cl          # size queries return a count, not a status
    lib.mumpy_last_error.argtypes = []
    lib.mumpy_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mumpy_last_error().decode("utf-8", "replace")
        raise MumpyError("%s failed (%d): %s" % (what or "libmumpy_b200 call", rc, msg))
