"""ctypes binding of libstedm_b200.so (the C ABI in include/stedm_b200.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry point raises, and
every non-zero return code becomes a RuntimeError carrying ``stedm_last_error()``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# STEDM_B200_LIB: another build of the same ABI (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("STEDM_B200_LIB") or os.path.join(_HERE, "libstedm_b200.so")

F32, BF16 = 0, 1

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class ConvDesc(C.Structure):
    """Mirror of ``struct stedm_conv_desc``."""
    _fields_ = [("x0", vp), ("x1", vp), ("weight", vp), ("bias", vp), ("emb", vp), ("residual", vp), ("out", vp),
                ("stats_out", vp), ("workspace", vp), ("workspace_bytes", C.c_int64), ("c0", C.c_int32), ("c1", C.c_int32), ("in_dtype", C.c_int32), ("batch", C.c_int32),
                ("in_h", C.c_int32), ("in_w", C.c_int32), ("x1_batch", C.c_int32), ("ksize", C.c_int32),
                ("stride", C.c_int32), ("upsample", C.c_int32), ("emb_stride", C.c_int32), ("res_dtype", C.c_int32),
                ("out_dtype", C.c_int32), ("out_nchw", C.c_int32), ("cout", C.c_int32), ("cout_store", C.c_int32), ("tap_mode", C.c_int32), ("phase", C.c_int32),
                ("act", C.c_int32), ("skip_c0", C.c_int32), ("skip_c1", C.c_int32), ("skip_x1_batch", C.c_int32),
                ("skip_x0", vp), ("skip_x1", vp), ("x0_pix_stride", C.c_int32), ("res_batch", C.c_int32),
                ("x1_pix_stride", C.c_int32), ("gn_cstride", C.c_int32), ("gn_c_off", C.c_int32), ("gn_silu", C.c_int32),
                ("gn_coef", vp)]


# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "stedm_abi_version": [],
    "stedm_last_error": [],
    "stedm_device_supported": [],
    "stedm_cfg_ddim_step": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, f32, f32, f32, vp],
    "stedm_gn_num_chunks": [i32, i32],
    "stedm_gn_stats": [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "stedm_gn_apply": [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp, vp, f32, i32, vp, i32, vp],
    "stedm_gn_apply_split": [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, f32, i32, i32, vp, vp, vp],
    "stedm_gn_fold_tiles": [vp, i32, i32, i64, i32, i32, vp, i32, i32, i64, i32, i32, i32, vp, vp, vp, f32, i32, vp, vp],
    "stedm_rows_add_emb": [vp, i64, vp, i32, vp, i32, i64, i32, i32, vp, vp],
    "stedm_conv_tc": [C.POINTER(ConvDesc), vp],
    "stedm_conv_tc_workspace_bytes": [C.POINTER(ConvDesc)],
    "stedm_conv_tc_plan": [C.POINTER(ConvDesc), C.POINTER(C.c_int32)],
    "stedm_conv_simt": [C.POINTER(ConvDesc), vp],
    "stedm_gemm_simt": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i64, i64, i64, i64,
                        i64, i64, f32, vp],
    "stedm_softmax_rows": [vp, vp, i32, i64, i32, f32, i32, vp],
    "stedm_attention_tc": [vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, f32, i64, i32, i32, i64, i64, i64, vp],
    "stedm_upsample_nearest2x": [vp, vp, i32, i32, i32, i32, i32, vp],
    "stedm_im2col_3x3_s2": [vp, vp, i32, i32, i32, i32, i32, vp],
    "stedm_pack_nchw_to_nhwc": [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp],
    "stedm_nhwc_to_nchw_f32": [vp, i32, vp, i32, i32, i32, vp],
    "stedm_timestep_embedding": [vp, vp, i32, i32, vp],
    "stedm_timestep_embedding_f32": [vp, vp, i32, i32, vp],
    "stedm_linear": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "stedm_vq_nearest": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "stedm_spatial_rescale": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "stedm_image_to_uint8": [vp, vp, i32, i32, i32, vp],
    "stedm_patch_embed_ln": [vp, vp, vp, vp, vp, f32, vp, vp, i32, i32, i32, i32, vp],
    "stedm_layernorm": [vp, i32, vp, vp, vp, f32, vp, vp, i64, i32, vp],
    "stedm_window_attention": [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "stedm_patch_merge_gather": [vp, vp, i32, i32, i32, i32, i32, vp],
    "stedm_ln_meanpool": [vp, vp, vp, f32, vp, i32, i32, i32, vp],
    "stedm_set_reduce": [vp, vp, i32, i32, i32, i32, vp],
    "stedm_geglu": [vp, vp, i32, i64, i32, vp],
    "stedm_spt_patchify": [vp, vp, i32, i32, i32, i32, vp],
    "stedm_svit_assemble": [vp, i32, vp, vp, vp, i32, i32, i32, i32, vp],
    "stedm_token_mean": [vp, vp, i32, i32, i32, i32, vp],
}
_RESTYPES = {"stedm_last_error": C.c_char_p, "stedm_conv_tc_workspace_bytes": C.c_longlong}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m stedm_b200.build` "
                           f"(there is no CPU or PyTorch fallback for the native kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.stedm_abi_version() != 3:
        raise RuntimeError("libstedm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return (load().stedm_last_error() or b"").decode()


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"stedm_b200 native call {what} failed (code {rc}): {last_error()}")
