"""Thin tensor-level wrappers over the C ABI (pointers + sizes in, freshly allocated torch tensors out)."""
import torch

from .. import _native as N


def _prep(t, dtype):
    N.require_cuda(t)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def rq_quantize(x, codebook, depth, want_list=True):
    """x [n,C], codebook [K,C] -> (quant_list [D,n,C] or None, codes [n,D] int64)"""
    x = _prep(x, torch.float32)
    cb = _prep(codebook, torch.float32)
    n, C = x.shape
    codes = torch.empty(n, depth, dtype=torch.int64, device=x.device)
    ql = torch.empty(depth, n, C, dtype=torch.float32, device=x.device) if want_list else None
    with torch.cuda.device(x.device):
        N.check(N.lib().rqb200_rq_quantize(N.ptr(x), N.ptr(cb), n, cb.shape[0], C, depth, N.ptr(codes), N.ptr(ql),
                                           None, N.stream_ptr(x.device)), "rq_quantize")
    N.launch_count["total"] += 1 if n else 0
    return ql, codes


def rq_embed(codes, codebook, summed):
    """codes [n,D] int64 -> [n,C] (summed over depth) or [n,D,C]"""
    codes = _prep(codes, torch.int64)
    cb = _prep(codebook, torch.float32)
    n, D = codes.shape
    K, C = cb.shape
    out = torch.empty((n, C) if summed else (n, D, C), dtype=torch.float32, device=codes.device)
    fn = N.lib().rqb200_rq_embed_sum if summed else N.lib().rqb200_rq_embed_depth
    with torch.cuda.device(codes.device):
        N.check(fn(N.ptr(codes), N.ptr(cb), n, D, K, C, N.ptr(out), N.stream_ptr(codes.device)), "rq_embed")
    N.launch_count["total"] += 1 if n else 0
    return out


def sample_logits(logits, temperature=1.0, top_k=None, top_p=None, q=None):
    logits = _prep(logits, torch.float32)
    B, V = logits.shape
    if q is not None:
        q = _prep(q, torch.float32)
    out = torch.empty(B, dtype=torch.int64, device=logits.device)
    k = 0 if top_k is None else int(top_k)
    p = 1.0 if top_p is None else float(top_p)
    with torch.cuda.device(logits.device):
        N.check(N.lib().rqb200_sample_logits(N.ptr(logits), N.ptr(q), B, V, float(temperature), k, p, N.ptr(out),
                                             N.stream_ptr(logits.device)), "sample_logits")
    N.launch_count["total"] += 1
    return out


def rq_soft(residual, codebook, temp=1.0, want_logits=False):
    """residual [n,C] -> softmax(-distances / temp) [n,K] (and optionally the logits -d/temp)"""
    r = _prep(residual, torch.float32)
    cb = _prep(codebook, torch.float32)
    n, C = r.shape
    K = cb.shape[0]
    soft = torch.empty(n, K, dtype=torch.float32, device=r.device)
    logits = torch.empty(n, K, dtype=torch.float32, device=r.device) if want_logits else None
    with torch.cuda.device(r.device):
        N.check(N.lib().rqb200_rq_soft_codes(N.ptr(r), N.ptr(cb), n, K, C, float(temp), N.ptr(soft), N.ptr(logits),
                                             N.stream_ptr(r.device)), "rq_soft_codes")
    N.launch_count["total"] += 1 if n else 0
    return (soft, logits) if want_logits else soft
