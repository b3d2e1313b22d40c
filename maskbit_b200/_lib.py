"""ctypes binding of libmaskbit_b200.so (include/maskbit_b200.h).

The library is the product: there is no Python / PyTorch / CPU fallback.  If the shared object is missing the
import of any compute entry point fails loudly with the build command.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MASKBIT_B200_LIB") or os.path.join(_HERE, "csrc", "libmaskbit_b200.so")   # env: A/B builds in tools/

MB_GENERATOR, MB_TOKENIZER = 0, 1


class MBConfig(ctypes.Structure):
    _fields_ = [("hidden_dim", ctypes.c_int), ("depth", ctypes.c_int), ("heads", ctypes.c_int), ("mlp_dim", ctypes.c_int),
                ("token_bits", ctypes.c_int), ("codebook_splits", ctypes.c_int), ("nclass", ctypes.c_int),
                ("seq_len", ctypes.c_int), ("use_prenorm", ctypes.c_int), ("dec_hidden_channels", ctypes.c_int),
                ("dec_channel_mult", ctypes.c_int * 8), ("dec_num_resolutions", ctypes.c_int),
                ("dec_num_res_blocks", ctypes.c_int), ("num_channels", ctypes.c_int), ("generator_cls", ctypes.c_int),
                ("enc_num_res_blocks", ctypes.c_int)]


class MBSelectArgs(ctypes.Structure):
    _fields_ = [("logits_c", ctypes.c_void_p), ("logits_u", ctypes.c_void_p), ("q", ctypes.c_void_p),
                ("gumbel", ctypes.c_void_p), ("tokens_in", ctypes.c_void_p), ("predicted", ctypes.c_void_p),
                ("tokens_out", ctypes.c_void_p), ("scale", ctypes.c_float), ("temperature", ctypes.c_float),
                ("randomize_temperature", ctypes.c_float), ("one_minus_progress", ctypes.c_float),
                ("mask_len", ctypes.c_float), ("B", ctypes.c_int), ("n", ctypes.c_int), ("splits", ctypes.c_int),
                ("V", ctypes.c_int), ("seq_stride", ctypes.c_int), ("mask_token", ctypes.c_int64),
                ("seed", ctypes.c_uint64), ("step", ctypes.c_uint32)]


class MBSampleArgs(ctypes.Structure):
    _fields_ = [("labels", ctypes.c_void_p), ("B", ctypes.c_int), ("num_steps", ctypes.c_int),
                ("use_guidance", ctypes.c_int), ("skip_zero_scale_uncond", ctypes.c_int),
                ("scale", ctypes.POINTER(ctypes.c_float)), ("temperature", ctypes.POINTER(ctypes.c_float)),
                ("one_minus_progress", ctypes.POINTER(ctypes.c_float)), ("mask_len", ctypes.POINTER(ctypes.c_float)),
                ("randomize_temperature", ctypes.c_float), ("q", ctypes.c_void_p), ("gumbel", ctypes.c_void_p),
                ("seed", ctypes.c_uint64), ("images", ctypes.c_void_p), ("trace", ctypes.c_void_p),
                ("final_tokens", ctypes.c_void_p)]


# every symbol include/maskbit_b200.h declares: (restype, argtypes)
_P, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
SYMBOLS = {
    "mb_create": (_I, [ctypes.POINTER(MBConfig), ctypes.POINTER(_P)]),
    "mb_destroy": (None, [_P]),
    "mb_last_error": (ctypes.c_char_p, []),
    "mb_version": (ctypes.c_char_p, []),
    "mb_set_tensor": (_I, [_P, _I, ctypes.c_char_p, _P, ctypes.POINTER(ctypes.c_int64), _I, _I]),
    "mb_finalize": (_I, [_P, _I]),
    "mb_generator_forward": (_I, [_P, _P, _I, _P, _I, _P, _I, _P, _P]),
    "mb_generator_forward_attn": (_I, [_P, _P, _I, _P, _I, _P, _I, _P, _P, _P]),
    "mb_select_step": (_I, [_P, ctypes.POINTER(MBSelectArgs), _P]),
    "mb_decode_tokens": (_I, [_P, _P, _I, _P, _P]),
    "mb_encode": (_I, [_P, _P, _I, _P, _P, _P]),
    "mb_combine_tokens": (_I, [_P, _P, _I, _P, _P]),
    "mb_postprocess_u8": (_I, [_P, _P, _I, _P, _P]),
    "mb_sample": (_I, [_P, ctypes.POINTER(MBSampleArgs), _P]),
    "mb_split_tokens": (_I, [_P, ctypes.c_int64, _I, _I, _P, _P]),
    "mb_mask_tokens": (_I, [_P, _P, _P, ctypes.c_int64, _P, _P, _I, _I, _P]),
    "mb_mlm_loss_scratch_bytes": (_I, []),
    "mb_mlm_loss": (_I, [_P, _P, _P, ctypes.c_int64, _I, _I, _F, _I, _P, _P, _P]),
    "mb_launch_count": (ctypes.c_int64, [_P]),
    "mb_profile_enable": (_I, [_P, _I]),
    "mb_profile_read": (_I, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), _I]),
    "mb_test_gemm": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "mb_test_attention": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mb_test_noise_transform": (_I, [_P, _P, _P, _P, _I, _P]),
    "mb_test_attention_trace": (_I, [_P]),
    "mb_test_gemm_trace": (_I, [_P]),
    "mb_test_gemm_ex": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
}

PROF_KINDS = ["embed_ln", "gemm_qkv", "attention", "gemm_out", "layernorm", "gemm_up", "gemm_down", "gemm_head", "select",
              "dec_conv", "dec_groupnorm", "dec_io"]

_lib = None


class MaskbitError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MaskbitError(
                f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `bash maskbit_b200/csrc/build.sh`. maskbit_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            if os.environ.get("MASKBIT_B200_LIB") and not hasattr(L, name):
                continue   # A/B builds of older sources (tools/build_variants.sh) may predate an entry point
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise MaskbitError(f"libmaskbit_b200 error {status}: {lib().mb_last_error().decode()}")


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
