"""ctypes binding of libhvit_sm100.so (C ABI declared in include/hvit.h)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libhvit_sm100.so")

MAX_STAGES = 8
MAX_LAYERS = 48
PREC_FP32 = 0
PREC_BF16 = 1
PREC_FP16 = 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2

_VP = C.c_void_p


class ModelCfg(C.Structure):
    _fields_ = [
        ("n_enc", C.c_int), ("enc_channels", C.c_int * MAX_STAGES), ("enc_pool", C.c_int * MAX_STAGES),
        ("embed_dim", C.c_int), ("num_heads", C.c_int), ("num_layers", C.c_int), ("mlp_hidden", C.c_int),
        ("patch_size", C.c_int),
        ("n_dec", C.c_int), ("dec_channels", C.c_int * MAX_STAGES), ("dec_up", C.c_int * MAX_STAGES),
        ("use_skip", C.c_int), ("precision", C.c_int), ("ln_eps", C.c_float),
    ]


class Weights(C.Structure):
    _fields_ = [
        ("stem_w", _VP), ("stem_scale", _VP), ("stem_shift", _VP),
        ("enc_w", _VP * MAX_STAGES), ("enc_scale", _VP * MAX_STAGES), ("enc_shift", _VP * MAX_STAGES),
        ("patch_w", _VP), ("patch_b", _VP), ("pos_embed", _VP), ("pos_len", C.c_int),
        ("ln1_g", _VP * MAX_LAYERS), ("ln1_b", _VP * MAX_LAYERS), ("ln2_g", _VP * MAX_LAYERS),
        ("ln2_b", _VP * MAX_LAYERS),
        ("qkv_w", _VP * MAX_LAYERS), ("qkv_b", _VP * MAX_LAYERS), ("proj_w", _VP * MAX_LAYERS),
        ("proj_b", _VP * MAX_LAYERS),
        ("fc1_w", _VP * MAX_LAYERS), ("fc1_b", _VP * MAX_LAYERS), ("fc2_w", _VP * MAX_LAYERS),
        ("fc2_b", _VP * MAX_LAYERS),
        ("lnf_g", _VP), ("lnf_b", _VP), ("tofm_w", _VP), ("tofm_b", _VP),
        ("skip_w", _VP * MAX_STAGES), ("skip_b", _VP * MAX_STAGES),
        ("dec_w", _VP * MAX_STAGES), ("dec_scale", _VP * MAX_STAGES), ("dec_shift", _VP * MAX_STAGES),
        ("head_w", _VP),
    ]


class RefWeights(C.Structure):
    """struct hvit_ref_weights: the reference state_dict as fp32 device pointers (input of hvit_pack_weights)."""
    _fields_ = [
        ("enc_conv_w", _VP * MAX_STAGES), ("enc_bn_w", _VP * MAX_STAGES), ("enc_bn_b", _VP * MAX_STAGES),
        ("enc_bn_mean", _VP * MAX_STAGES), ("enc_bn_var", _VP * MAX_STAGES),
        ("patch_w", _VP), ("patch_b", _VP), ("pos_embed", _VP), ("pos_len", C.c_int),
        ("ln1_w", _VP * MAX_LAYERS), ("ln1_b", _VP * MAX_LAYERS), ("ln2_w", _VP * MAX_LAYERS),
        ("ln2_b", _VP * MAX_LAYERS),
        ("qkv_w", _VP * MAX_LAYERS), ("qkv_b", _VP * MAX_LAYERS), ("proj_w", _VP * MAX_LAYERS),
        ("proj_b", _VP * MAX_LAYERS),
        ("fc1_w", _VP * MAX_LAYERS), ("fc1_b", _VP * MAX_LAYERS), ("fc2_w", _VP * MAX_LAYERS),
        ("fc2_b", _VP * MAX_LAYERS),
        ("lnf_w", _VP), ("lnf_b", _VP), ("tofm_w", _VP), ("tofm_b", _VP),
        ("dec_conv_w", _VP * MAX_STAGES), ("dec_bn_w", _VP * MAX_STAGES), ("dec_bn_b", _VP * MAX_STAGES),
        ("dec_bn_mean", _VP * MAX_STAGES), ("dec_bn_var", _VP * MAX_STAGES),
        ("skip_w", _VP * MAX_STAGES), ("skip_b", _VP * MAX_STAGES),
    ]


# name -> (restype, argtypes); every symbol include/hvit.h declares
_I, _F, _SZ = C.c_int, C.c_float, C.c_size_t
SYMBOLS = {
    "hvit_last_error": (C.c_char_p, []),
    "hvit_version": (_I, []),
    "hvit_device_ok": (_I, []),
    "hvit_packed_weights_bytes": (_SZ, [C.POINTER(ModelCfg), _I]),
    "hvit_pack_weights": (_I, [C.POINTER(ModelCfg), C.POINTER(RefWeights), _VP, _SZ, C.POINTER(Weights), _VP]),
    "hvit_workspace_bytes": (_SZ, [C.POINTER(ModelCfg), _I, _I, _I, _I]),
    "hvit_plan_create": (_I, [C.POINTER(ModelCfg), C.POINTER(Weights), _I, _I, _I, _I, _VP, _SZ, _VP, C.POINTER(_VP)]),
    "hvit_plan_destroy": (None, [_VP]),
    "hvit_plan_set_debug": (_I, [_VP, _I]),
    "hvit_forward": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "hvit_enhance": (_I, [_VP, _VP, _VP, _I, _VP]),
    "hvit_enhance_varlen": (_I, [_VP, _VP, _VP, _VP, _I, _VP]),
    "hvit_varlen_min_samples": (_I, [_VP]),
    "hvit_metrics_scratch_bytes": (_SZ, [_I, _I]),
    "hvit_metrics": (_I, [_VP, _VP, _I, _I, _VP, _VP, _SZ, _VP, _VP]),
    "hvit_spec_loss": (_I, [_VP, _VP, _I, C.c_longlong, _I, _VP, _VP]),
    "hvit_plan_buffer": (_I, [_VP, C.c_char_p, C.POINTER(_SZ), C.POINTER(_I * 4), C.POINTER(_I)]),
    "hvit_plan_launch_count": (_I, [_VP, _I]),
    "hvit_plan_tokens": (_I, [_VP, C.POINTER(_I), C.POINTER(_I)]),
    "hvit_plan_num_steps": (_I, [_VP, _I]),
    "hvit_plan_step_info": (_I, [_VP, _I, _I, C.c_char_p, _I, C.c_char_p, _I, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I)]),
    "hvit_enhance_profiled": (_I, [_VP, _VP, _VP, _I, _VP, C.POINTER(C.c_float), _I]),
    "hvit_gemm_16": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _I, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "hvit_linear_ln_producer_16": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "hvit_linear_ln_consumer_16": (_I, [_VP, _VP, _I, _VP, _VP, _VP, _VP, _F, _I, _VP, _I, _I, _I, _I, _I, _VP, _VP, _VP]),
    "hvit_rowstats_16": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "hvit_gemm_f32": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _I, _VP, _I, _I, _I, _I, _VP]),
    "hvit_conv3x3_16": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "hvit_conv3x3_f32": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP, _I, _I, _I, _I, _I, _VP]),
    "hvit_stem_16": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "hvit_head_16": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "hvit_patch_embed_16": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _VP, _I, _I, _VP, _I, _VP]),
    "hvit_skip_concat_16": (_I, [_VP, _I, _I, _I, _I, _VP, _VP, _I, _VP, _I, _I, _I, _I, _VP, _I, _VP]),
    "hvit_attention_16": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "hvit_attention_f32": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP]),
    "hvit_layernorm": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _F, _VP]),
    "hvit_stft": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "hvit_istft": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP]),
}

_lib = None


class HvitError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HvitError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). hvit_b200 has no CPU / PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "hvit") -> None:
    if status != 0:
        msg = load().hvit_last_error()
        raise HvitError(f"{what} failed (status {status}): {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
