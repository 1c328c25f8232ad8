"""ctypes binding of the C ABI in include/rqb200.h (csrc/librqb200.so).  Fails loudly when the library is absent."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RQB200_LIB", os.path.join(os.path.dirname(_HERE), "csrc", "librqb200.so"))

OK, EINVAL, ECUDA, ENODEV, EWORKSPACE, ESTATE = 0, -1, -2, -3, -4, -5
F32, BF16, F16 = 0, 1, 2
MODE_EXACT, MODE_FAST = 0, 1
AR_NO_GRAPH, AR_NO_PDL, AR_TRACE, AR_L2_PREFETCH, AR_SHALLOW_RING, AR_SEQUENTIAL_PREFILL, AR_BATCHED_DEEP_RING, AR_BATCHED_STREAMER = 1, 2, 4, 8, 16, 32, 64, 128
AR_ATTN_ONE_WARP, AR_TRACE_WEIGHTS, AR_NO_PARAM_PREFETCH = 256, 512, 1024
_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

c_f32p, c_i64p, c_vp = C.c_void_p, C.c_void_p, C.c_void_p


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wqkv", "wproj", "w1", "w2", "bqkv", "bproj", "b1", "b2",
                                          "ln1_w", "ln1_b", "ln2_w", "ln2_b")]


class ArConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("embed_dim", "n_head", "n_body", "n_head_layers", "vocab", "H", "W", "D",
                                         "vocab_cond", "cond_len", "code_dim", "codebook_size", "mode", "weight_dtype",
                                         "flags", "split_qkv", "split_proj", "split_fc1", "split_fc2")]


class ArWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("pos_emb_cond", "pos_emb_hw", "pos_emb_d", "cond_emb", "w_in", "w_head", "w_cls",
                                          "b_in", "b_head", "b_cls", "cls_ln_w", "cls_ln_b", "codebook")] + \
               [("body", C.POINTER(BlockWeights)), ("head", C.POINTER(BlockWeights))] + \
               [(n, C.c_void_p) for n in ("w_ccls", "b_ccls", "ccls_ln_w", "ccls_ln_b")]


class VaeConfig(C.Structure):
    _fields_ = [("ch", C.c_int32), ("n_levels", C.c_int32), ("ch_mult", C.c_int32 * 8), ("num_res_blocks", C.c_int32),
                ("n_attn_res", C.c_int32), ("attn_resolutions", C.c_int32 * 8), ("resolution", C.c_int32),
                ("z_channels", C.c_int32), ("embed_dim", C.c_int32), ("in_channels", C.c_int32), ("out_ch", C.c_int32),
                ("codebook_size", C.c_int32), ("depth", C.c_int32), ("mode", C.c_int32)]


_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Loads librqb200.so once.  No fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError("rqb200: native library not found at %s -- build it with "
                          "rq-vae-transformer_b200/csrc/build.sh (or __graft_entry__.build()); there is no CPU fallback"
                          % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.rqb200_last_error.restype = C.c_char_p
    L.rqb200_version.restype = C.c_int
    L.rqb200_device_count.restype = C.c_int
    L.rqb200_rq_quantize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    L.rqb200_rq_soft_codes.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    L.rqb200_rq_embed_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.rqb200_rq_embed_depth.argtypes = L.rqb200_rq_embed_sum.argtypes
    L.rqb200_sample_logits.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p,
                                       C.c_void_p]
    L.rqb200_dbg_rows_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64,
                                       C.c_int, C.c_int, C.c_void_p]
    L.rqb200_dbg_tma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float),
                                      C.POINTER(C.c_float)]
    L.rqb200_dbg_rq_quantize.argtypes = [C.c_int] + L.rqb200_rq_quantize.argtypes
    L.rqb200_dbg_sample_logits.argtypes = [C.c_int] + L.rqb200_sample_logits.argtypes
    L.rqb200_ar_create.restype = C.c_void_p
    L.rqb200_ar_create.argtypes = [C.POINTER(ArConfig), C.POINTER(ArWeights)]
    L.rqb200_ar_destroy.argtypes = [C.c_void_p]
    L.rqb200_ar_destroy.restype = None
    L.rqb200_ar_workspace_bytes.restype = C.c_size_t
    L.rqb200_ar_workspace_bytes.argtypes = [C.c_void_p, C.c_int]
    L.rqb200_ar_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.rqb200_ar_sample_span.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.rqb200_ar_forward_workspace_bytes.restype = C.c_size_t
    L.rqb200_ar_forward_workspace_bytes.argtypes = [C.c_void_p, C.c_int]
    L.rqb200_ar_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]
    L.rqb200_ar_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    L.rqb200_ar_last_launches.restype = C.c_int64
    L.rqb200_ar_last_launches.argtypes = [C.c_void_p]
    L.rqb200_vae_create.restype = C.c_void_p
    L.rqb200_vae_create.argtypes = [C.POINTER(VaeConfig)]
    L.rqb200_vae_destroy.argtypes = [C.c_void_p]
    L.rqb200_vae_destroy.restype = None
    L.rqb200_vae_set_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64]
    L.rqb200_vae_finalize.argtypes = [C.c_void_p]
    L.rqb200_vae_workspace_bytes.restype = C.c_size_t
    L.rqb200_vae_workspace_bytes.argtypes = [C.c_void_p, C.c_int]
    for fn in (L.rqb200_vae_decode, L.rqb200_vae_decode_code, L.rqb200_vae_encode):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.rqb200_vae_last_launches.restype = C.c_int64
    L.rqb200_vae_last_launches.argtypes = [C.c_void_p]
    L.rqb200_dbg_gemm_tc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.rqb200_dbg_conv_tc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.rqb200_dbg_chain.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                   C.POINTER(C.c_float)]
    L.rqb200_dbg_chain2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_float)]
    _lib = L
    return L


EXPORTS = ["rqb200_last_error", "rqb200_version", "rqb200_device_count", "rqb200_rq_quantize", "rqb200_rq_embed_sum",
           "rqb200_rq_embed_depth", "rqb200_rq_soft_codes", "rqb200_sample_logits", "rqb200_ar_create", "rqb200_ar_destroy",
           "rqb200_ar_workspace_bytes", "rqb200_ar_sample", "rqb200_ar_sample_span", "rqb200_ar_forward",
           "rqb200_ar_forward_workspace_bytes", "rqb200_ar_trace", "rqb200_ar_last_launches", "rqb200_vae_create",
           "rqb200_vae_destroy", "rqb200_vae_set_tensor", "rqb200_vae_finalize", "rqb200_vae_workspace_bytes",
           "rqb200_vae_decode", "rqb200_vae_decode_code", "rqb200_vae_encode", "rqb200_vae_last_launches",
           "rqb200_dbg_gemm_tc", "rqb200_dbg_conv_tc", "rqb200_dbg_chain", "rqb200_dbg_chain2", "rqb200_dbg_rq_quantize",
           "rqb200_dbg_sample_logits", "rqb200_dbg_tma_rate", "rqb200_dbg_rows_gemm"]


def check(rc, what=""):
    if rc != 0:
        raise NativeError("rqb200 %s failed (%d): %s" % (what, rc, lib().rqb200_last_error().decode()))


def require_cuda(*tensors):
    """the product path is CUDA-only: refuse CPU tensors instead of silently computing elsewhere"""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NativeError("rqb200: tensor on %s -- this engine has no CPU path; move the model and inputs to a "
                              "CUDA device" % t.device)


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dtype_code(t):
    return _DT[t.dtype]


def default_precision():
    """'exact' (fp32 FFMA, bit-exact-indices gate) or 'fast' (fp16/bf16 tcgen05).  RQB200_PRECISION overrides."""
    return os.environ.get("RQB200_PRECISION", "auto")


def fast_dtype():
    """16-bit operand format of the fast AR tier: fp16 (the reference's autocast class) unless RQB200_FAST_DTYPE=bf16"""
    return torch.bfloat16 if os.environ.get("RQB200_FAST_DTYPE", "fp16").lower() in ("bf16", "bfloat16") else torch.float16


def ar_engine_options():
    """engine-creation options of the fast AR tier, read ONCE per engine from the environment (diagnostics / experiments)"""
    env = os.environ.get
    flags = 0
    flags |= AR_NO_GRAPH if env("RQB200_NO_GRAPH", "0") == "1" else 0
    flags |= AR_NO_PDL if env("RQB200_NO_PDL", "0") == "1" else 0
    flags |= AR_TRACE if env("RQB200_TRACE", "0") in ("1", "2") else 0
    flags |= AR_TRACE_WEIGHTS if env("RQB200_TRACE", "0") == "2" else 0
    flags |= AR_L2_PREFETCH if env("RQB200_GEMM_L2PF", "0") == "1" else 0
    flags |= AR_SHALLOW_RING if env("RQB200_GEMM_SHALLOW", "0") == "1" else 0
    flags |= AR_SEQUENTIAL_PREFILL if env("RQB200_SEQ_PREFILL", "0") == "1" else 0
    flags |= AR_BATCHED_DEEP_RING if env("RQB200_BATCHED_DEEP", "0") == "1" else 0
    flags |= AR_BATCHED_STREAMER if env("RQB200_BATCHED_STREAMER", "0") == "1" else 0
    flags |= AR_ATTN_ONE_WARP if env("RQB200_ATTN_ONE_WARP", "0") == "1" else 0
    flags |= AR_NO_PARAM_PREFETCH if env("RQB200_NO_PARAM_PREFETCH", "0") == "1" else 0
    return {"flags": flags,
            "splits": [int(env("RQB200_SPLIT_" + k, "0")) for k in ("QKV", "PROJ", "FC1", "FC2")]}


def param_fingerprint(module):
    """identity + in-place-write counter of every parameter / buffer.  The native engines cache packed weight copies; they are
    rebuilt when this changes (nested load_state_dict through a wrapper, EMA updates, optimizer steps, .data swaps)."""
    return hash(tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers())))


launch_count = {"total": 0}      # kernels launched by our library through this binding (bench.py's gpu_launches)
