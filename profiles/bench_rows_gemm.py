"""The large-M GEMM of the batched passes, shape by shape (1.4B widths): TFLOP/s of launch_rows_gemm_tc.
RQB200_ROWS_GEMM_1CTA=1 selects the single-CTA persistent kernel (128 x 256 tiles) for every shape; default: CTA pairs (256 x 256).
usage: python profiles/bench_rows_gemm.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae import _native as N  # noqa: E402

L = N.lib()
dev = "cuda"
which = "single CTA (128 x 256)" if os.environ.get("RQB200_ROWS_GEMM_1CTA", "0") == "1" else "CTA pairs (256 x 256)"
print("rows GEMM, %s, fp16" % which)
tot_f, tot_t = 0.0, 0.0
for M, N_out, K, mode in ((4096, 4608, 1536, "h16"), (4096, 1536, 1536, "f32+res"), (4096, 6144, 1536, "h16 gelu"), (4096, 1536, 6144, "f32+res"),
                          (16384, 4608, 1536, "h16"), (16384, 6144, 1536, "h16 gelu"), (16384, 1536, 6144, "f32+res"), (16384, 16384, 1536, "f32")):
    g = torch.Generator().manual_seed(1)
    W = (torch.randn(N_out, K, generator=g) / K ** 0.5).half().to(dev)
    X = torch.randn(M + 256, K, generator=g).half().to(dev)
    bias = torch.randn(N_out, generator=g).to(dev)
    of = torch.zeros(M, N_out, device=dev) if mode.startswith("f32") else None
    oh = torch.zeros(M, N_out, dtype=torch.float16, device=dev) if of is None else None
    res = of if "res" in mode else None

    def run():
        N.check(L.rqb200_dbg_rows_gemm(N.ptr(X), N.ptr(W), N.ptr(bias), N.ptr(res), N.ptr(of), N.ptr(oh), 1 if "gelu" in mode else 0, 0,
                                       M, N_out, K, N.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * M * N_out * K
    tot_f += fl
    tot_t += ms
    print("M %5d  N %5d  K %4d  %-9s : %7.3f ms  %6.1f TFLOP/s" % (M, N_out, K, mode, ms, fl / ms / 1e9), flush=True)
print("all shapes: %.1f TFLOP/s" % (tot_f / tot_t / 1e9))
