"""Teacher-forced RQTransformer.forward on the fast tier: the large-M tcgen05 GEMM path.
Times model(codes, amp=True) and reports achieved TFLOP/s against the measured dense bf16/fp16 peak.
usage: python profiles/bench_forward.py [model] [B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "in1400m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
ar, vae, dd = bench.build_models(name, dev, "fast", tiny_vae=True)
E, nh, nb, nhl, V, bs, vc, cl = bench.MODELS[name][:8]
H, W, D = bs
codes = torch.randint(0, V, (B, H, W, D), device=dev)
cond = torch.randint(0, max(vc, 1), (B, cl), device=dev)
for it in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ar(codes, model_aux=vae, cond=cond, amp=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
Tb, Th = cl + H * W - 1, H * W * D
flops = 2.0 * B * (Tb * nb * 12 * E * E + Th * (nhl * 12 * E * E + E * V))
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
pk = float(peaks.get("bf16_tflops_sustained", 1454.1))
print("%s B=%d: forward %.2f ms, %.1f GFLOP of GEMM work -> %.1f TFLOP/s = %.1f %% of the measured sustained dense 16-bit peak (%.0f TFLOP/s); "
      "%d launches" % (name, B, ms, flops / 1e9, flops / ms / 1e9, 100 * flops / ms / 1e9 / pk, pk, ar.last_launches))
