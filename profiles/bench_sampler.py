"""micro-benchmark of the fused sampler kernel.  usage: python profiles/bench_sampler.py [B] [V]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae.models import _bind as nb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
V = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
logits = torch.randn(B, V, device="cuda") * 0.6
q = torch.empty(B, V, device="cuda").exponential_(1)
for (k, p) in ((1024, 1.0), (1024, 0.95), (None, 1.0), (1, 1.0)):
    for _ in range(5):
        nb.sample_logits(logits, 1.0, k, p, q=q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 200
    for _ in range(n):
        nb.sample_logits(logits, 1.0, k, p, q=q)
    e1.record()
    torch.cuda.synchronize()
    print("sampler B=%d V=%d top_k=%s top_p=%s: %.1f us per call (incl. launch)" % (B, V, k, p, e0.elapsed_time(e1) * 1e3 / n))
