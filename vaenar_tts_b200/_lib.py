"""ctypes binding of libvaenar_sm100.so (include/vaenar_b200.h).  There is no fallback: if the shared
library is missing or a call fails, an exception is raised."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAENAR_LIB") or os.path.join(_HERE, "libvaenar_sm100.so")   # VAENAR_LIB: tuning builds only


class VaenarError(RuntimeError):
    pass


class HParamsStruct(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "vocab_size", "embd_dim", "enc_n_conv", "enc_hidden", "enc_conv_kernel", "enc_n_blk", "enc_att_dim",
        "enc_heads", "enc_ffn",
        "dec_nblk", "dec_att_dim", "dec_heads", "dec_ffn", "post_n_conv", "post_filters", "post_kernel",
        "posterior_pre_hidden", "posterior_nblk", "posterior_att_dim", "posterior_heads", "posterior_ffn",
        "prior_n_blk", "prior_n_tblk", "prior_att_dim", "prior_heads", "prior_ffn",
        "latent_dim", "out_dim", "max_reduction_factor", "final_reduction_factor")] + [(n, c_float) for n in (
        "mel_text_len_ratio", "enc_pre_drop_rate", "enc_pos_drop_rate", "posterior_pre_drop_rate",
        "posterior_pos_drop_rate", "post_drop_rate")]


class TrainOpts(ctypes.Structure):
    _fields_ = [("n_masks", c_int32), ("masks", POINTER(c_void_p)), ("seed", c_uint64), ("update_bn_stats", c_int32)]


_P = c_void_p
_SIGS = {
    "vaenar_last_error": (c_char_p, []),
    "vaenar_abi_version": (c_int, []),
    "vaenar_create": (c_int, [POINTER(HParamsStruct), POINTER(c_void_p)]),
    "vaenar_destroy": (c_int, [_P]),
    "vaenar_num_params": (c_int, [_P]),
    "vaenar_param_name": (c_char_p, [_P, c_int]),
    "vaenar_param_ndim": (c_int, [_P, c_int]),
    "vaenar_param_dim": (c_int64, [_P, c_int, c_int]),
    "vaenar_param_offset": (c_int64, [_P, c_int]),
    "vaenar_param_trainable": (c_int, [_P, c_int]),
    "vaenar_param_floats": (c_int64, [_P]),
    "vaenar_packed_bytes": (c_int64, [_P]),
    "vaenar_workspace_bytes": (c_int64, [_P, c_int, c_int, c_int, c_int]),
    "vaenar_pack_weights": (c_int, [_P, _P, _P, _P]),
    "vaenar_pack_weights_async": (c_int, [_P, _P, _P, _P]),
    "vaenar_text_encoder_fwd": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, c_int, c_int, c_float, _P, _P]),
    "vaenar_length_predictor_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P]),
    "vaenar_prior_sample": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "vaenar_prior_log_probability": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "vaenar_posterior_fwd": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                     _P, _P, _P]),
    "vaenar_posterior_params": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "vaenar_decoder_fwd": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "vaenar_inference": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                 _P]),
    "vaenar_elbo_fwd": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P,
                                _P, _P, _P, _P, _P]),
    "vaenar_elbo_fwd_train": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                      POINTER(TrainOpts), _P, _P, _P, _P, _P, _P]),
    "vaenar_init": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, POINTER(TrainOpts), _P, _P, _P]),
    "vaenar_prior_init": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "vaenar_train_workspace_bytes": (c_int64, [_P, c_int, c_int, c_int, c_int]),
    "vaenar_train_step_grads": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                        POINTER(TrainOpts), c_float, c_float, c_float, _P, _P, _P, _P]),
    "vaenar_trainable_mask": (c_int, [_P, _P]),
    "vaenar_adam_step": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_float, c_float, c_float, c_float, c_float, _P, _P]),
    "vaenar_grad_nonfinite": (c_int, [_P, c_int64, _P, _P]),
    "vaenar_enable_peer_access": (c_int, [c_int]),
    "vaenar_ipc_export": (c_int, [_P, _P, POINTER(c_int64)]),
    "vaenar_ipc_open": (c_int, [_P, POINTER(c_void_p)]),
    "vaenar_ipc_close": (c_int, [_P]),
    "vaenar_adam_shard_floats": (c_int64, [c_int64, c_int]),
    "vaenar_adam_step_sharded": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int64, c_float, c_float, c_float, c_float,
                                         c_float, _P, _P]),
    "vaenar_crc32c": (ctypes.c_uint32, [_P, c_int64, ctypes.c_uint32]),
    "vaenar_randn": (c_int, [_P, c_int64, c_uint64, c_uint64, c_float, _P]),
    "vaenar_launch_count": (ctypes.c_long, []),
    "vaenar_profile_enable": (c_int, [c_int]),
    "vaenar_profile_report": (c_char_p, []),
    "vaenar_debug_gemm_timestamps": (c_int, [_P]),
    "vaenar_debug_gemm_launches": (c_char_p, []),
    "vaenar_test_dense": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P,
                                  c_int64, _P]),
    "vaenar_test_conv1d": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    "vaenar_test_attention": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_int64, _P]),
    "vaenar_test_attention_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P,
                                          c_int64, _P]),
    "vaenar_test_wgrad": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    "vaenar_xblk_stack_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "vaenar_set_fused": (c_int, [c_int]),
    "vaenar_set_train_fused": (c_int, [c_int]),
    "vaenar_set_gemm_occ2": (c_int, [c_int]),
    "vaenar_debug_xrow_timestamps": (c_int, [_P]),
    "vaenar_griffin_lim_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "vaenar_mel_to_linear": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_float, _P,
                                     _P]),
    "vaenar_griffin_lim": (c_int, [_P, _P, _P, c_uint64, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int64, _P, c_int64,
                                   _P]),
    "vaenar_inv_preemphasis": (c_int, [_P, c_int64, _P, c_int, c_int, c_int, c_int, ctypes.c_double, _P, c_int64, _P]),
    "vaenar_wav_to_int16": (c_int, [_P, c_int64, _P, c_int, c_int, c_int, c_int, _P, c_int64, _P, _P]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def load():
    """Load the shared library (once) and declare every prototype of include/vaenar_b200.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VaenarError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The B200 path has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise VaenarError(load().vaenar_last_error().decode(errors="replace"))
