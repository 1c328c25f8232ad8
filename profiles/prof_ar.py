"""profiling driver: a few AR positions of the 1.4B-shaped transformer (same widths/depths, smaller code grid) so that
an ncu launch list stays short.  usage: python profiles/prof_ar.py [B] [H] [W] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prec = sys.argv[4] if len(sys.argv) > 4 else "fast"
name = os.environ.get("PROF_MODEL", "in1400m")
m = list(bench.MODELS[name])
m[5] = (H, W, 4)
bench.MODELS["prof"] = tuple(m)
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
ar, vae, dd = bench.build_models("prof", dev, prec, tiny_vae=True)
aux = vae          # the RQ-VAE itself supplies the shared codebook (model_aux.get_code_emb_with_depth path)
part = torch.zeros(B, H, W, 4, dtype=torch.long, device=dev)
cond = torch.randint(0, 1000, (B, 1), device=dev)
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    codes = ar.sample(part, model_aux=aux, cond=cond, top_k=1024, amp=(prec == "fast"))
    e1.record()
    torch.cuda.synchronize()
    print("iter %d: %.3f ms per position (%d positions, B=%d)" % (it, e0.elapsed_time(e1) / (H * W), H * W, B))
