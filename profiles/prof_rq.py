"""P1 measurement: fused RQ search (N = B*8*8 vectors, K codewords, D = 4) -- achieved FP32 FLOP/s and the HBM figure the
north star asks for.  usage: python profiles/prof_rq.py [B] [K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae.models import _bind as nb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
N, C, D = B * 64, 256, 4
dev = "cuda"
cb = torch.randn(K, C, device=dev)
xs = [torch.randn(N, C, device=dev) * 0.2 for _ in range(8)]
for x in xs[:3]:
    nb.rq_quantize(x, cb, D, want_list=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 16
for i in range(reps):
    nb.rq_quantize(xs[i % 8], cb, D, want_list=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flop = 2.0 * N * K * C * D
hbm = N * C * 4 + K * C * 4 + N * D * 8 + D * N * C * 4
print("rq_quantize N=%d K=%d D=%d: %.3f ms  %.1f TFLOP/s fp32 (FFMA peak ~72 TFLOP/s)  algorithmic HBM %.1f MB -> %.1f GB/s "
      "(%.2f%% of 6578 GB/s: compute-bound by construction, SURVEY finding 4)" % (N, K, D, ms, flop / ms / 1e9, hbm / 1e6, hbm / ms / 1e6, hbm / ms / 1e6 / 6578 * 100))
