"""micro-benchmark of the tcgen05 weight-streaming GEMM alone (cold weights: a ring of distinct weight matrices larger
than L2).  usage: python profiles/bench_gemm.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae import _native as N  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = N.lib()
dev = "cuda"
E = 1536
shapes = [("qkv", 3 * E, E), ("proj", E, E), ("fc1", 4 * E, E), ("fc2", E, 4 * E), ("cls", 16384, E)]
print("B=%d" % B)
for name, n_out, k in shapes:
    nw = max(4, int(600e6 // (n_out * k * 2)))
    Ws = [torch.randn(n_out, k, device=dev).to(torch.bfloat16) for _ in range(nw)]
    X = torch.randn(B, k, device=dev).to(torch.bfloat16)
    for splits in (1, 2, 3, 4, 6, 8, 12):
        if splits > k // 64 or (n_out // 128) * splits > 1200:
            continue
        part = torch.empty(splits, B, n_out, device=dev)
        out = torch.empty(B, n_out, device=dev)
        st = N.stream_ptr()

        def run(i):
            if splits == 1:
                L.rqb200_dbg_gemm_tc(N.ptr(Ws[i % nw]), N.ptr(X), None, None, N.ptr(out), 0, 0, None, n_out, k, B, 1, 1, st)
            else:
                L.rqb200_dbg_gemm_tc(N.ptr(Ws[i % nw]), N.ptr(X), None, None, None, 0, 0, N.ptr(part), n_out, k, B, splits, 1, st)
        for i in range(nw):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3 * nw
        for i in range(n):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / n
        gb = n_out * k * 2 / 1e9
        print("%-5s N=%5d K=%5d splits=%2d ctas=%4d : %7.2f us  %7.1f GB/s" % (name, n_out, k, splits, n_out // 128 * splits, us, gb / (us * 1e-6)))
