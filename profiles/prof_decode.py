"""profiling driver: the ImageNet RQ-VAE decoder (fast tier) on a batch of random codes.
usage: python profiles/prof_decode.py [B] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
m = list(bench.MODELS["tiny"])
m[4] = 16384
bench.MODELS["prof"] = tuple(m)
ar, vae, dd = bench.build_models("prof", dev, "fast")
codes = torch.randint(0, 16384, (B, 8, 8, 4), device=dev)
for it in range(iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pix = vae.decode_code(codes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("iter %d: decode %.2f ms for %d images (%.0f TFLOP/s algorithmic at 250 GFLOP/img; x3 tensor work in split-fp16 mode)" % (it, ms, B, 0.25 * B / (ms * 1e-3)))
