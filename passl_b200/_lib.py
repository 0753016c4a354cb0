"""ctypes binding of the C-ABI library (include/passl_b200.h).

The library is the product: there is no Python / CPU fallback.  Importing a kernel wrapper without the built
``libpassl_b200.so`` raises immediately (run ``python __graft_entry__.py`` to build it).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpassl_b200.so")

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

# name -> (restype, [argtypes]); mirrors include/passl_b200.h one to one
SIGNATURES = {
    "passl_b200_version": (c_int, []),
    "passl_b200_launch_count": (c_ll, []),
    "passl_b200_launch_counter_add": (c_ll, [c_ll]),
    "passl_b200_gemm_stats_rows": (c_int, []),
    "passl_b200_gemm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll,
                                     c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p,
                                     c_void_p]),
    "passl_b200_gemm_bf16_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll,
                                        c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p,
                                        c_void_p, c_int, c_void_p, c_void_p]),
    "passl_b200_conv2d_fwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 9 +
                                   [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "passl_b200_wgrad_halo_mode": (c_int, [c_int]),
    "passl_b200_conv2d_fwd_rect_bf16": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 11 + [c_void_p, c_int, c_void_p, c_void_p]),
    "passl_b200_conv2d_wgrad_rect_bf16": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 12 + [c_void_p]),
    "passl_b200_stem_pack_input": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "passl_b200_stem_pack_weight": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "passl_b200_stem_unpack_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "passl_b200_conv2d_dgrad_workspace_bytes": (c_ll, [c_int] * 4),
    "passl_b200_conv2d_dgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]),
    "passl_b200_conv2d_wgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p]),
    "passl_b200_simce_workspace_bytes": (c_ll, [c_int, c_int]),
    "passl_b200_simce_fwd_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                         c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll,
                                         c_void_p]),
    "passl_b200_simce_bwd_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                         c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll,
                                         c_void_p]),
    "passl_b200_infonce_tc_workspace_bytes": (c_ll, [c_int, c_int, c_int]),
    "passl_b200_infonce_tc_set_debug": (c_int, [c_void_p]),
    "passl_b200_infonce_tc_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_int,
                                          c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "passl_b200_infonce_tc_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_int,
                                          c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "passl_b200_infonce_tc_fwd_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_uint, c_void_p, c_void_p, c_float,
                                               c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "passl_b200_infonce_tc_bwd_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_uint, c_void_p, c_void_p, c_float,
                                               c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "passl_b200_peer_publish_keys_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, ctypes.c_uint, c_void_p,
                                                  c_void_p]),
    "passl_b200_ntxent_workspace_bytes": (c_ll, [c_int]),
    "passl_b200_ntxent_co2_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_ll, c_void_p]),
    "passl_b200_ntxent_co2_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "passl_b200_l2norm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "passl_b200_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                      c_void_p]),
    "passl_b200_queue_enqueue": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "passl_b200_ema_update": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_ll, c_void_p]),
    "passl_b200_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "passl_b200_cast_bf16_to_f32": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "passl_b200_bn_reduce_blocks": (c_int, [c_ll, c_int]),
    "passl_b200_bn_stats": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "passl_b200_bn_finalize": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_ll, c_float, c_float, c_int, c_void_p]),
    "passl_b200_bn_bwd_finalize": (c_int, [c_void_p, c_int] + [c_void_p] * 6 + [c_ll, c_void_p, c_int, c_void_p]),
    "passl_b200_bn_global_affine": (c_int, [c_void_p] * 8 + [c_float, c_int, c_void_p]),
    "passl_b200_axpy_f32": (c_int, [c_void_p, c_void_p, c_float, c_ll, c_void_p]),
    "passl_b200_bn_apply": (c_int, [c_void_p] * 6 + [c_ll, c_int, c_int, c_void_p]),
    "passl_b200_bn_apply_mask": (c_int, [c_void_p] * 6 + [c_ll, c_int, c_int, c_void_p]),
    "passl_b200_bn_bwd_reduce": (c_int, [c_void_p] * 6 + [c_ll, c_int, c_int, c_void_p]),
    "passl_b200_bn_bwd_apply": (c_int, [c_void_p] * 6 + [c_ll, c_int, c_int, c_void_p]),
    "passl_b200_mae_random_masking": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "passl_b200_token_assemble_fwd": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "passl_b200_token_assemble_bwd": (c_int, [c_void_p] * 6 + [c_int] * 6 + [c_void_p]),
    "passl_b200_mae_loss_workspace_bytes": (c_ll, [c_int, c_int]),
    "passl_b200_mae_loss_fwd": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_float, c_void_p, c_void_p]),
    "passl_b200_mae_loss_bwd": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_float, c_void_p]),
    "passl_b200_embedding_fwd": (c_int, [c_void_p] * 4 + [c_ll, c_int, c_int, c_int, c_void_p]),
    "passl_b200_embedding_bwd": (c_int, [c_void_p] * 4 + [c_ll, c_int, c_int, c_int, c_void_p]),
    "passl_b200_eot_gather_fwd": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "passl_b200_eot_gather_bwd": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p]),
    "passl_b200_clip_ce_workspace_bytes": (c_ll, [c_int]),
    "passl_b200_clip_ce_fwd": (c_int, [c_void_p] * 3 + [c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "passl_b200_clip_ce_bwd": (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_ll, c_void_p]),
    "passl_b200_rows_ce_fwd": (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_ll, c_void_p]),
    "passl_b200_rows_ce_bwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_void_p]),
    "passl_b200_peer_buffer_create": (c_int, [c_ll, c_void_p, c_void_p]),
    "passl_b200_peer_buffer_open": (c_int, [c_void_p, c_void_p]),
    "passl_b200_peer_buffer_close": (c_int, [c_void_p]),
    "passl_b200_peer_buffer_destroy": (c_int, [c_void_p]),
    "passl_b200_peer_allgather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, ctypes.c_uint, c_void_p]),
    "passl_b200_peer_reduce_scatter_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, ctypes.c_uint, c_void_p]),
    "passl_b200_umma_probe": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "passl_b200_resample_kmax": (c_int, [c_int] * 3),
    "passl_b200_resized_crop_workspace_bytes": (c_ll, [c_int] * 4),
    "passl_b200_resized_crop_u8": (c_int, [c_void_p] * 8 + [c_ll] + [c_int] * 5 + [c_void_p]),
    "passl_b200_views_finalize_f32": (c_int, [c_void_p] * 4 + [c_int] * 2 + [ctypes.c_double] + [c_void_p] * 3),
    "passl_b200_color_jitter_u8": (c_int, [c_void_p] * 4 + [c_ll] + [c_int] * 3 + [c_void_p]),
    "passl_b200_gaussian_blur_workspace_bytes": (c_ll, [c_int] * 2),
    "passl_b200_gaussian_blur_u8": (c_int, [c_void_p] * 4 + [c_ll] + [c_int] * 3 + [c_void_p]),
    "passl_b200_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "passl_b200_attention_bwd": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "passl_b200_layernorm_fwd": (c_int, [c_void_p] * 6 + [c_ll, c_int, c_float, c_void_p]),
    "passl_b200_layernorm_bwd_blocks": (c_int, [c_ll]),
    "passl_b200_layernorm_bwd": (c_int, [c_void_p] * 8 + [c_ll, c_int, c_void_p]),
    "passl_b200_im2col_nchw_f32": (c_int, [c_void_p, c_void_p] + [c_int] * 9 + [c_void_p]),
    "passl_b200_maxpool3x3s2_fwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 4 + [c_void_p]),
    "passl_b200_bn_relu_maxpool3x3s2_fwd": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "passl_b200_maxpool3x3s2_bwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 4 + [c_void_p]),
    "passl_b200_avgpool_fwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 3 + [c_void_p]),
    "passl_b200_avgpool_bwd": (c_int, [c_void_p, c_void_p] + [c_int] * 3 + [c_void_p]),
    "passl_b200_sgd_momentum": (c_int, [c_void_p] * 4 + [c_float] * 4 + [c_void_p, c_ll, c_void_p]),
    "passl_b200_lars_momentum": (c_int, [c_void_p] * 7 + [c_int] + [c_float] * 5 + [c_void_p, c_ll, c_void_p]),
    "passl_b200_adamw": (c_int, [c_void_p] * 8 + [c_float] * 4 + [c_int, c_float, c_void_p, c_ll, c_void_p]),
    "passl_b200_grad_norm_finite": (c_int, [c_void_p, c_ll, c_float, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p]),
}

_ERRORS = {-1: "bad argument / contract violation", -2: "unsupported configuration", -3: "TMA tensor-map encode failed",
           -4: "workspace too small"}


class PasslB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load libpassl_b200.so and attach prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PasslB200Error(
            "passl_b200: %s is missing — the CUDA extension is mandatory (no CPU fallback). "
            "Build it with `python __graft_entry__.py`." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what):
    if code == 0:
        return
    if code < 0:
        raise PasslB200Error("%s: %s (code %d)" % (what, _ERRORS.get(code, "error"), code))
    raise PasslB200Error("%s: CUDA error %d" % (what, code))
