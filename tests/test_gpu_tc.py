"""tcgen05 tier, kernel level: the weight-streaming GEMM against an fp32 matmul of the same bf16-rounded operands."""
import pytest
import torch

from rqvae import _native as N

pytestmark = pytest.mark.gpu
DEV = "cuda"


def gelu(x):
    return torch.nn.functional.gelu(x)


@pytest.mark.parametrize("N_out,K,B,splits", [(128, 64, 16, 1), (256, 128, 1, 1), (384, 128, 3, 1), (1536, 1536, 64, 1),
                                               (4608, 1536, 64, 1), (1536, 6144, 64, 6), (1536, 1536, 8, 4),
                                               (16384, 1536, 64, 1), (6144, 1536, 200, 1), (1536, 1536, 33, 24),
                                               (2048, 1024, 128, 2)])
def test_gemm_tc_matches_fp32_matmul(N_out, K, B, splits):
    g = torch.Generator().manual_seed(N_out + K + B)
    W = (torch.randn(N_out, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    X = torch.randn(B, K, generator=g).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N_out, generator=g).to(DEV)
    R = torch.randn(B, N_out, generator=g).to(DEV)
    ref = X.float() @ W.float().t()
    L = N.lib()
    st = N.stream_ptr()
    if splits == 1:
        out = torch.empty(B, N_out, device=DEV)
        N.check(L.rqb200_dbg_gemm_tc(N.ptr(W), N.ptr(X), N.ptr(bias), N.ptr(R), N.ptr(out), 0, 0, None, N_out, K, B, 1, st))
        torch.cuda.synchronize()
        # fp32 accumulate of exact bf16 products: only the summation order differs
        torch.testing.assert_close(out, ref + bias + R, rtol=1e-4, atol=1e-4)
        outb = torch.empty(B, N_out, device=DEV, dtype=torch.bfloat16)
        N.check(L.rqb200_dbg_gemm_tc(N.ptr(W), N.ptr(X), N.ptr(bias), None, N.ptr(outb), 1, 1, None, N_out, K, B, 1, st))
        torch.cuda.synchronize()
        torch.testing.assert_close(outb.float(), gelu(ref + bias).to(torch.bfloat16).float(), rtol=2e-2, atol=2e-2)
    else:
        part = torch.full((splits, B, N_out), float("nan"), device=DEV)
        N.check(L.rqb200_dbg_gemm_tc(N.ptr(W), N.ptr(X), None, None, None, 0, 0, N.ptr(part), N_out, K, B, splits, st))
        torch.cuda.synchronize()
        torch.testing.assert_close(part.sum(0), ref, rtol=1e-4, atol=1e-4)
