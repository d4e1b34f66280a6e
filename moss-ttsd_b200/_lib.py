"""ctypes binding of libmtts.so (the C-ABI declared in include/mtts.h).

There is deliberately no fallback: if the shared library cannot be loaded the import of any op raises,
so a GPU run can never silently execute something other than the CUDA path.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmtts.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "mtts.h")

_lib = None
_lock = threading.Lock()

c_void_p, c_int, c_ll, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t, ctypes.c_float
c_double = ctypes.c_double
c_u64 = ctypes.c_ulonglong

class SamplerConfig(ctypes.Structure):
    """Mirror of `mtts_sampler_config` (include/mtts.h)."""
    _fields_ = [
        ("channels", c_int),
        ("vocab", c_int * 8), ("logit_offset", c_int * 8), ("do_sample", c_int * 8),
        ("has_rep", c_int * 8), ("rep_penalty", c_float * 8),
        ("has_temp", c_int * 8), ("temperature", c_float * 8),
        ("top_k", c_int * 8),
        ("has_top_p", c_int * 8), ("top_p", c_float * 8),
        ("seen_offset_words", c_int * 8), ("seen_words_per_row", c_int),
        ("pad_token", c_int), ("eos_mask_token", c_int),
    ]


_cfg_p = ctypes.POINTER(SamplerConfig)


class LMLayer(ctypes.Structure):
    """Mirror of `mtts_lm_layer`: ten device pointers per decoder layer."""
    _fields_ = [(n, c_void_p) for n in ("wqkv", "wo", "wgu", "wd", "ln1", "ln2", "q_norm", "k_norm", "k_pool", "v_pool")]


class DecodeMegaArgs(ctypes.Structure):
    """Mirror of `mtts_decode_mega_args` (include/mtts.h)."""
    _fields_ = [
        ("layers", c_void_p), ("num_layers", c_int),
        ("hidden", c_int), ("intermediate", c_int), ("num_q_heads", c_int), ("num_kv_heads", c_int), ("head_dim", c_int),
        ("heads", c_void_p), ("vpad", c_int),
        ("final_norm", c_void_p), ("inv_freq", c_void_p), ("positions", c_void_p), ("block_table", c_void_p),
        ("max_pages", c_int), ("page_size", c_int), ("num_pages", c_int),
        ("x", c_void_p),
        ("logits", c_void_p), ("ld_logits", c_ll),
        ("B", c_int), ("nsplit", c_int), ("eps", c_float),
        ("workspace", c_void_p), ("workspace_bytes", c_ll),
        ("err_flag", c_void_p), ("profile_cycles", c_void_p),
    ]

# name -> (restype, argtypes). Kept in the same order as include/mtts.h; tests/test_abi.py checks that every
# symbol the header declares is listed here and exported by the .so.
SIGNATURES = {
    "mtts_last_error": (ctypes.c_char_p, []),
    "mtts_version": (c_int, []),
    "mtts_init": (c_int, []),
    "mtts_launch_count": (c_ll, []),
    "mtts_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mtts_gemm": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_size_t, c_void_p]),
    "mtts_gemm_splitk_splits": (c_int, [c_int, c_int, c_int]),
    "mtts_gemm_splitk_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mtts_gemm_splitk": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mtts_splitk_reduce": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "mtts_splitk_reduce_rmsnorm": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_float,
                                           c_void_p]),
    "mtts_gemm_simt": (c_int, [c_void_p, c_int, c_ll, c_ll, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "mtts_rvq_codebook_norms": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mtts_rvq_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "mtts_rvq_decode": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mtts_embed_sum8": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int), c_int, c_void_p,
                                c_void_p, c_void_p]),
    "mtts_rmsnorm": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_void_p]),
    "mtts_qknorm_rope_kvappend": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_float, c_void_p, c_void_p]),
    "mtts_gqa_attention_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mtts_gqa_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                                   c_void_p]),
    "mtts_sampler_init_history": (c_int, [c_void_p, c_int, c_int, c_ll, _cfg_p, c_void_p, c_void_p]),
    "mtts_sample8_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mtts_sample8": (c_int, [c_void_p, c_ll, c_int, _cfg_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "mtts_sample8_rows": (c_int, [c_void_p, c_ll, c_int, _cfg_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "mtts_heads8_sample_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mtts_heads8_sample_fused": (c_int, [_cfg_p, c_int]),
    "mtts_heads8_sample": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, _cfg_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mtts_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_float, c_void_p, c_int, c_void_p]),
    "mtts_layernorm_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_float, c_void_p, c_int, c_void_p]),
    "mtts_dwconv7_ln_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                    c_void_p]),
    "mtts_mha_varlen": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_mha_varlen_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_mha_varlen_fp32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_gqa_prefill_tc": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_mha_varlen_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_rows_prefix_copy": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mtts_split_tf32x3": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_int, c_void_p]),
    "mtts_dwconv7_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                c_void_p]),
    "mtts_convt_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_im2col": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_stft_frames": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_power_spectrum": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p]),
    "mtts_logmel_finish": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "mtts_istft_spec": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_void_p]),
    "mtts_istft_ola": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mtts_istft_head_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_ll]),
    "mtts_istft_head": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_int,
                                c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mtts_add_rows_mod": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "mtts_gqa_decode_fused": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                      c_size_t, c_void_p, c_void_p]),
    "mtts_gqa_decode_fused_splitk": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                             c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                             c_size_t, c_void_p, c_void_p]),
    "mtts_decode_mega_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "mtts_decode_mega_workspace_bytes": (c_ll, [c_int, c_int]),
    "mtts_decode_mega": (c_int, [ctypes.POINTER(DecodeMegaArgs), c_void_p]),
    "mtts_delay_step": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, _cfg_p, c_void_p]),
    "mtts_delay_step_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, _cfg_p,
                                     c_void_p]),
}


class MttsError(RuntimeError):
    pass


def header_symbols(path: str = HEADER_PATH):
    """Function names declared in include/mtts.h."""
    with open(path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mtts_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True):
    """Load libmtts.so, building it in-tree with nvcc first if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise MttsError(f"{LIB_PATH} is missing; run `python moss-ttsd_b200/build.py`")
            from . import build as _build
            _build.build(verbose=False)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/.so mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        msg = load().mtts_last_error()
        raise MttsError(f"libmtts error {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
