"""ctypes binding of libmmad.so (include/mmad.h).  No fallback: if the shared library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmad.so")

MMAD_MAX_LAYERS = 16
PREC = {"fp32": 0, "f16x3": 1, "f16": 2, "f16f8": 3}


class MmadError(RuntimeError):
    pass


class Desc(C.Structure):
    _fields_ = [("n_enc", C.c_int), ("n_dec", C.c_int),
                ("enc_widths", C.c_int * (MMAD_MAX_LAYERS + 1)),
                ("dec_widths", C.c_int * (MMAD_MAX_LAYERS + 1)),
                ("lrelu_slope", C.c_float), ("bn_eps", C.c_float), ("precision", C.c_int)]


class TrainLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("W", "b", "gamma", "beta", "run_mean", "run_var", "num_batches_tracked", "gW", "gb", "ggamma", "gbeta")]


class FeatureWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("conv1r_w", "conv1r_b", "conv2r_w", "conv2r_b", "conv3r_w", "conv3r_b",
                 "conv1d_w", "conv1d_b", "conv2d_w", "conv2d_b", "conv3d_w", "conv3d_b",
                 "conv1l_w", "conv1l_b", "conv2l_w", "conv2l_b")]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p)

_vp, _i, _sz, _ll, _f = C.c_void_p, C.c_int, C.c_size_t, C.c_longlong, C.c_float

# name -> (restype, argtypes); every symbol include/mmad.h declares
SIGNATURES = {
    "mmad_last_error": (C.c_char_p, []),
    "mmad_version": (_i, []),
    "mmad_create": (_i, [C.POINTER(Desc), C.POINTER(_vp)]),
    "mmad_destroy": (_i, [_vp]),
    "mmad_set_precision": (_i, [_vp, _i]),
    "mmad_set_option": (_i, [_vp, C.c_char_p, C.c_double]),
    "mmad_set_layer": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmad_workspace_bytes": (_sz, [_vp, _i]),
    "mmad_ae_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "mmad_recon_loss": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "mmad_score": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mmad_score_host": (_i, [_vp, _vp, _i, _ll, _i, _i, _vp, _vp, _vp]),
    "mmad_stream_input": (_i, [_vp, _i, _i, C.POINTER(C.POINTER(C.c_float)), C.POINTER(_i)]),
    "mmad_concat_width": (_i, [_vp, _i, _i]),
    "mmad_nap_accumulate_sum": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "mmad_nap_accumulate_gram": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "mmad_nap_set_fit": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mmad_nap_rotate_stats": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "mmad_nap_set_standardizer": (_i, [_vp, _vp, _vp, _vp]),
    "mmad_nap_set_structure": (_i, [_vp, _i]),
    "mmad_fc_layer_forward": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _i, _vp]),
    "mmad_sq_diff_sum": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "mmad_scale_unless_one": (_i, [_vp, _ll, _vp, _vp]),
    "mmad_row_mean_sq": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "mmad_vib_reparam": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mmad_normalizer_workspace_bytes": (_sz, [_i]),
    "mmad_col_stats": (_i, [_vp, _i, _ll, _i, _vp, _vp, _vp, _sz, _vp]),
    "mmad_gram_accumulate": (_i, [_vp, _i, _ll, _i, _vp, _vp, _vp, _sz, _vp]),
    "mmad_rotate": (_i, [_vp, _i, _ll, _i, _vp, _vp, _i, _vp, _i, _vp, _sz, _vp]),
    "mmad_standardize": (_i, [_vp, _i, _ll, _i, _vp, _vp, _vp, _i, _vp]),
    "mmad_tri_pack": (_i, [_vp, _i, _vp, _vp]),
    "mmad_tri_unpack": (_i, [_vp, _i, _vp, _vp]),
    "mmad_metric_workspace_bytes": (_sz, [_ll]),
    "mmad_auc_roc": (_i, [_vp, _vp, _ll, C.POINTER(C.c_double), _vp, _sz, _vp]),
    "mmad_auc_prc": (_i, [_vp, _vp, _ll, C.POINTER(C.c_double), _vp, _sz, _vp]),
    "mmad_quantile": (_i, [_vp, _ll, _f, C.POINTER(C.c_float), _vp, _sz, _vp]),
    "mmad_confusion": (_i, [_vp, _vp, _ll, _f, _i, C.POINTER(C.c_longlong), _vp, _sz, _vp]),
    "mmad_comm_unique_id": (_i, [_vp]),
    "mmad_comm_init": (_i, [_vp, _vp, _i, _i]),
    "mmad_comm_destroy": (_i, [_vp]),
    "mmad_comm_world": (_i, [_vp]),
    "mmad_comm_set_grad_allreduce": (_i, [_vp, _i]),
    "mmad_comm_allreduce_f32": (_i, [_vp, _vp, _ll, _vp]),
    "mmad_comm_allreduce_f64": (_i, [_vp, _vp, _ll, _vp]),
    "mmad_peer_create": (_i, [_vp, _vp]),
    "mmad_peer_open": (_i, [_vp, _vp, _i, _i]),
    "mmad_peer_close": (_i, [_vp]),
    "mmad_peer_allreduce_f64": (_i, [_vp, _vp, _ll, _vp]),
    "mmad_peer_grad_alloc": (_i, [_vp, _ll, _vp, _vp]),
    "mmad_peer_grad_open": (_i, [_vp, _vp]),
    "mmad_multisensory_width": (_i, [_i, _i, _i, _i]),
    "mmad_multisensory_forward": (_i, [_vp, _vp, _vp, _vp, _i, C.POINTER(FeatureWeights), C.POINTER(C.c_float), _vp, _i, _vp]),
    "mmad_profile_begin": (_i, [_vp]),
    "mmad_profile_end": (_i, [_vp, C.POINTER(C.c_double)]),
    "mmad_launch_count": (C.c_ulonglong, []),
    "mmad_train_workspace_bytes": (_sz, [_vp, _i]),
    "mmad_train_loss": (_i, [_vp, _vp]),
    "mmad_train_fwd_bwd": (_i, [_vp, _vp, _i, _i, _ll, C.POINTER(TrainLayer), C.POINTER(TrainLayer), _vp, _f, _f,
                                _vp, _vp, _sz, ALLREDUCE_FN, _vp, _vp]),
    "mmad_adam_step": (_i, [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                            C.POINTER(_ll), _i, _f, _f, _f, _f, _f, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libmmad.so (built in-tree by ``csrc/build.py`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmadError(f"{LIB_PATH} is missing: run `python -m icra2021_multimodal_ad_b200.csrc.build` "
                            "(there is no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MmadError(f"libmmad error {rc}: {lib().mmad_last_error().decode()}")
