"""BASELINE config 1 on the GPU (encode -> quantize(D=4) -> decode, FFHQ RQ-VAE K=2048) and the rFID / code-extraction path
(get_codes) at batch 2 and 64, exact and fast tier.  usage: python profiles/prof_codes.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
_, vae, dd = bench.build_models("ffhq355m", dev, "fast")


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


for prec in ("fast", "exact"):
    vae.precision = prec
    for B in (2, 64):
        x = torch.randn(B, 3, 256, 256, device=dev)
        t_enc, z = timed(lambda: vae.encode(x))
        t_q, (ql, codes) = timed(lambda: vae.quantizer.quantize(z))
        t_dec, pix = timed(lambda: vae.decode_code(codes))
        t_all, _ = timed(lambda: vae(x))
        print("%-5s B=%2d: encode %8.2f ms  quantize %6.3f ms  decode %8.2f ms  | forward (config 1 round trip) %8.2f ms = %7.1f images/s"
              % (prec, B, t_enc, t_q, t_dec, t_all, B / t_all * 1e3), flush=True)
