"""ctypes binding of libhonerf_b200.so (the C ABI declared in include/honerf_b200.h).

There is no fallback: if the shared library is missing this module raises at import, and every
compute entry point needs a CUDA device.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhonerf_b200.so")

HN_MAX_LAYERS = 12
HN_SIMT_FP32, HN_TC_TF32, HN_TC_TF32X3, HN_TC_BF16X3, HN_TC_MIXED16 = 0, 1, 2, 3, 4
HN_WS_SDF_ONLY, HN_WS_FWD, HN_WS_BWD = 0, 1, 2


class hn_mlp_t(Structure):
    _fields_ = [("n_layers", c_int32),
                ("in_dim", c_int32 * HN_MAX_LAYERS),
                ("out_dim", c_int32 * HN_MAX_LAYERS),
                ("ld", c_int32 * HN_MAX_LAYERS),
                ("W", c_void_p * HN_MAX_LAYERS),
                ("b", c_void_p * HN_MAX_LAYERS),
                ("WT", c_void_p * HN_MAX_LAYERS),
                ("ldT", c_int32 * HN_MAX_LAYERS),
                ("chain", c_void_p),
                ("chain_bytes", c_int64)]


class hn_wn_job_t(Structure):
    _fields_ = [("v", c_void_p), ("g", c_void_p), ("dW", c_void_p), ("W", c_void_p), ("WT", c_void_p),
                ("dv", c_void_p), ("dg", c_void_p),
                ("out_dim", c_int32), ("in_dim", c_int32), ("ld", c_int32), ("ldT", c_int32),
                ("gap_at", c_int32), ("gap", c_int32), ("post_scale", c_float), ("pad_", c_int32)]


class hn_mlp_grad_t(Structure):
    _fields_ = [("dW", c_void_p * HN_MAX_LAYERS),
                ("db", c_void_p * HN_MAX_LAYERS)]


class _LazyLib:
    """The shared library, loaded on first use: host-only modules (honerf_b200.dist, the PackedMLP layout helpers) import
    on a machine that has not built it; the first C-ABI call on such a machine raises a clear ImportError, and a stale
    library (a symbol of include/honerf_b200.h missing) says "rebuild" instead of an AttributeError deep inside ctypes.
    There is still no CPU or PyTorch fallback."""

    def __init__(self):
        object.__setattr__(self, "_cdll", None)

    def _load(self):
        if self._cdll is None:
            if not os.path.isfile(LIB_PATH):
                raise ImportError(
                    "honerf_b200: %s is missing. Build it with `python ho-nerf_b200/build.py` (nvcc, sm_100a); "
                    "there is no CPU or PyTorch fallback." % LIB_PATH)
            cdll = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                try:
                    fn = getattr(cdll, name)
                except AttributeError:
                    raise ImportError("honerf_b200: %s is stale (no symbol %s): rebuild it with "
                                      "`python ho-nerf_b200/build.py --force`" % (LIB_PATH, name)) from None
                fn.restype = res
                fn.argtypes = args
            object.__setattr__(self, "_cdll", cdll)
        return self._cdll

    def __getattr__(self, name):
        return getattr(self._load(), name)


lib = _LazyLib()

P = c_void_p
_mlp_p = POINTER(hn_mlp_t)
_grad_p = POINTER(hn_mlp_grad_t)

# name -> (restype, argtypes).  Mirrors include/honerf_b200.h one to one.
PROTOTYPES = {
    "hn_last_error": (ctypes.c_char_p, []),
    "hn_version": (c_int, []),
    "hn_launch_count": (c_int64, []),
    "hn_timing_enable": (c_int, [c_int]),
    "hn_timing_reset": (c_int, []),
    "hn_timing_collect": (c_int, [POINTER(ctypes.c_double), POINTER(c_int64)]),
    "hn_timing_collect_tags": (c_int, [POINTER(ctypes.c_double), POINTER(c_int64), c_int]),
    "hn_wn_pack": (c_int, [P, P, c_int, c_int, c_int, c_float, P, P, c_int, P]),
    "hn_wn_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_float, P, P, P]),
    "hn_wn_pack_batch": (c_int, [POINTER(hn_wn_job_t), c_int, P]),
    "hn_wn_bwd_batch": (c_int, [POINTER(hn_wn_job_t), c_int, P]),
    "hn_mlp_bx3_bytes": (c_int64, [_mlp_p]),
    "hn_mlp_bx3_pack": (c_int, [_mlp_p, P, c_int64, P]),
    "hn_sdf_hand_chain_bytes": (c_int64, [_mlp_p]),
    "hn_sdf_hand_chain_pack": (c_int, [_mlp_p, P, c_int64, P]),
    "hn_adam_flat": (c_int, [P, P, P, P, c_int64, P, P, P] + [ctypes.c_double] * 6 + [P]),
    "hn_peer_block_bytes": (c_int64, [c_int64, c_int]),
    "hn_peer_alloc": (c_int, [c_int64, P, P]),
    "hn_peer_open": (c_int, [P, P]),
    "hn_peer_close": (c_int, [P]),
    "hn_peer_free": (c_int, [P]),
    "hn_peer_adam_flat": (c_int, [P, P, P, c_int64, P, c_int, c_int, P, P, c_int] + [P, P] + [ctypes.c_double] * 6 + [P]),
    "hn_wn_pack_gap": (c_int, [P, P, c_int, c_int, c_int, c_float, c_int, c_int, P, P, c_int, P]),
    "hn_wn_bwd_gap": (c_int, [P, P, P, c_int, c_int, c_int, c_float, c_int, c_int, P, P, P]),
    "hn_sdf_hand_stash_floats": (c_int64, [c_int64]),
    "hn_sdf_hand_ws_floats": (c_int64, [c_int64, c_int]),
    "hn_sdf_hand_sdf": (c_int, [_mlp_p, P, P, P, c_int64, c_int64, P, P, c_int64, c_int, P]),
    "hn_sdf_hand_fwd": (c_int, [_mlp_p, P, P, P, c_int64, c_int64, P, P, c_int64, P, P, c_int64, P, c_int64,
                                c_int, P]),
    "hn_sdf_hand_fwd_render": (c_int, [_mlp_p, P, P, P, c_int64, c_int64, P, P, c_int64, P, P, c_int64, P, c_int64,
                                c_int, P]),
    "hn_sdf_hand_bwd": (c_int, [_mlp_p, P, P, P, c_int64, c_int64, P, P, P, c_int64, P, P, c_int64, P, P, P,
                                _grad_p, P, c_int64, c_int, P]),
    "hn_color_hand_stash_floats": (c_int64, [c_int64]),
    "hn_color_hand_ws_floats": (c_int64, [c_int64, c_int]),
    "hn_color_hand_fwd": (c_int, [_mlp_p, P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, c_int, P]),
    "hn_color_hand_fwd_render": (c_int, [_mlp_p, P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, c_int, P]),
    "hn_color_hand_chain_bytes": (c_int64, [_mlp_p]),
    "hn_color_hand_chain_pack": (c_int, [_mlp_p, P, c_int64, P]),
    "hn_color_hand_bwd": (c_int, [_mlp_p, c_int64, P, P, P, P, c_int64, P, c_int64, P, _grad_p, P, c_int64,
                                  c_int, P]),
    "hn_chain_set_prof": (c_int, [P]),
    "hn_chain_set_stagger": (c_int, [c_int, c_int]),
    "hn_sdf_obj_chain_bytes": (c_int64, []),
    "hn_sdf_obj_chain_pack": (c_int, [_mlp_p, P, c_int64, P]),
    "hn_sdf_obj_stash_floats": (c_int64, [c_int64]),
    "hn_sdf_obj_ws_floats": (c_int64, [c_int64, c_int]),
    "hn_sdf_obj_grid": (c_int, [_mlp_p, P, c_int, P, c_int, P, c_int, c_float, P, P]),
    "hn_sdf_obj_sdf": (c_int, [_mlp_p, P, c_int64, c_float, P, P, c_int64, c_int, P]),
    "hn_sdf_obj_fwd": (c_int, [_mlp_p, P, c_int64, c_float, P, P, c_int64, P, P, c_int64, P, c_int64,
                               c_int, P]),
    "hn_sdf_obj_bwd": (c_int, [_mlp_p, c_int64, c_float, P, P, P, c_int64, P, P, _grad_p, P, c_int64,
                               c_int, P]),
    "hn_color_obj_chain_bytes": (c_int64, []),
    "hn_color_obj_chain_pack": (c_int, [_mlp_p, P, c_int64, P]),
    "hn_color_obj_stash_floats": (c_int64, [c_int64]),
    "hn_color_obj_ws_floats": (c_int64, [c_int64, c_int]),
    "hn_color_obj_fwd": (c_int, [_mlp_p, P, P, P, c_int64, P, c_int64, P, P, c_int64, c_int, P]),
    "hn_color_obj_bwd": (c_int, [_mlp_p, c_int64, P, P, P, P, P, P, c_int64, P, _grad_p, P, c_int64,
                                 c_int, P]),
    "hn_dw_test": (c_int, [P, c_int64, c_int, c_int, P, c_int64, c_int, c_int, P, P, c_int64, P, c_int64, P, P, c_int64, P]),
    "hn_mc_set_tables": (c_int, [P, P, P]),
    "hn_mc_classify": (c_int, [P, c_int, c_int, c_int, c_float, P, P, P]),
    "hn_mc_emit": (c_int, [P, c_int, c_int, c_int, c_float, P, P, P, P, P, P, P]),
    "hn_outside_points": (c_int, [P, P, P, c_float, c_int64, c_int, P, P, P, P]),
    "hn_outside_composite_fwd": (c_int, [P, P, P, P, c_int64, c_int, P, P, P, P, P]),
    "hn_outside_composite_bwd": (c_int, [P, P, P, P, P, P, c_int64, c_int, P, P, P, P, P, P, P]),
    "hn_dw16_test": (c_int, [P, c_int, P, c_int, P, P, c_int64, P, c_int64, P, P, c_int64, P, c_int64, P]),
    "hn_dw16_set_debug": (c_int, [c_int]),
    "hn_chain16_set_debug": (c_int, [P]),
    "hn_ray_points": (c_int, [P, P, P, c_int64, c_int, P, P]),
    "hn_mid_points": (c_int, [P, P, P, c_int64, c_int, c_float, P, P, P, P]),
    "hn_mid_points_bwd": (c_int, [P, P, P, P, c_int64, c_int, P, P, P]),
    "hn_rays_to_local": (c_int, [P, P, P, P, c_int64, P, P, P]),
    "hn_rays_to_local_bwd": (c_int, [P, P, P, P, P, P, c_int64, P, P, P, P, P]),
    "hn_up_sample": (c_int, [P, P, P, c_int64, c_int, c_int, c_float, P, P]),
    "hn_inverse_cdf": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, P, P]),
    "hn_merge_sorted": (c_int, [P, c_int, P, c_int, c_int64, P, P, P, P, c_int64, P, P]),
    "hn_sort_rows": (c_int, [P, c_int64, c_int, P, P, P]),
    "hn_neus_composite_fwd": (c_int, [P, P, P, P, P, P, c_int64, c_int, c_int, P, P, P, P, P, P, P, P]),
    "hn_neus_composite_bwd": (c_int, [P, P, P, P, P, P, P, c_int64, c_int, c_int, P, P, P, P, P, P, P,
                                      P, P, P]),
    "hn_neus_alpha_fwd": (c_int, [P, P, P, P, P, c_int64, c_int, P, P, P]),
    "hn_neus_alpha_bwd": (c_int, [P, P, P, P, P, c_int64, c_int, P, P, P, P, P, P, P]),
    "hn_fit_composite_fwd": (c_int, [P, P, P, P, c_int64, c_int, P, P, P, P]),
    "hn_fit_composite_bwd": (c_int, [P, P, P, P, P, c_int64, c_int, P, P, P, P, P, P, P]),
    "hn_rays_from_ndc": (c_int, [P, P, c_int64, c_int64, P, P, P]),
    "hn_rays_ndc_grid": (c_int, [P, P, c_int, c_int, P, c_int64, c_int64, P, P, P]),
    "hn_nn_select": (c_int, [P, P, P, c_int, c_int, P, P, P]),
    "hn_loss_ws_floats": (c_int64, []),
    "hn_render_loss_fwd": (c_int, [P, P, P, P, P, c_int64, c_float, P, c_float, c_float, c_float, P, P, P]),
    "hn_render_loss_bwd": (c_int, [P, P, P, P, P, P, c_int64, c_float, c_float, c_float, P, P, P, P]),
    "hn_interaction_loss_fwd": (c_int, [P, c_int64, P, c_int64, c_int64, c_float, c_float, c_float, P, P, P]),
    "hn_interaction_loss_bwd": (c_int, [P, P, c_int64, P, c_int64, P, c_int64, c_float, c_float, c_float, P, P, P]),
}

# the tcgen05 bring-up self-test GEMMs: a separate library (csrc/selftest/), loaded by the tests only
SELFTEST_LIB_PATH = os.path.join(HERE, "libhonerf_b200_selftest.so")
SELFTEST_PROTOTYPES = {
    "hn_tc_gemm_test": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "hn_tc_gemm_ts_test": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "hn_gemm_test": (c_int, [c_int, c_int, c_int, c_int, c_int, P, c_int64, P, c_int64, P, P, c_int64, P]),
}


def load_selftest():
    if not os.path.isfile(SELFTEST_LIB_PATH):
        raise ImportError("honerf_b200: %s is missing (python ho-nerf_b200/build.py)" % SELFTEST_LIB_PATH)
    cdll = ctypes.CDLL(SELFTEST_LIB_PATH)
    for name, (res, args) in SELFTEST_PROTOTYPES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    cdll.hn_last_error.restype = ctypes.c_char_p
    return cdll


class HonerfError(RuntimeError):
    pass


def check(status, what):
    if status != 0:
        msg = lib.hn_last_error()
        raise HonerfError("%s failed (%d): %s" % (what, status, msg.decode() if msg else "?"))


def launch_count():
    return int(lib.hn_launch_count())
