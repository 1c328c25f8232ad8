"""experiment driver: time the AR stage of one model under several sets of RQB200_* environment switches in ONE process
(the switches are read when the native engine is created; the engine is dropped between settings).
usage: python profiles/exp_env.py "K=V,K=V" "K=V" ...   ("" = defaults).  env: PROF_MODEL, EXP_B, EXP_H, EXP_W"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402

B = int(os.environ.get("EXP_B", 64))
H = int(os.environ.get("EXP_H", 8))
W = int(os.environ.get("EXP_W", 8))
name = os.environ.get("PROF_MODEL", "in1400m")
m = list(bench.MODELS[name])
m[5] = (H, W, 4)
bench.MODELS["prof"] = tuple(m)
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
ar, vae, dd = bench.build_models("prof", dev, "fast", tiny_vae=True)
part = torch.zeros(B, H, W, 4, dtype=torch.long, device=dev)
cond = torch.randint(0, 1000, (B, 1), device=dev)
ref = None
for setting in (sys.argv[1:] or [""]):
    kv = [s.split("=", 1) for s in setting.split(",") if s]
    for k, v in kv:
        os.environ[k] = v
    ar._invalidate_native()
    best = 1e30
    try:
        for it in range(3):
            torch.manual_seed(1234)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            codes = ar.sample(part, model_aux=vae, cond=cond, top_k=1024, amp=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        same = "ref" if ref is None else ("same" if torch.equal(codes, ref) else "DIFF %d" % int((codes != ref).sum()))
        if ref is None:
            ref = codes.clone()
        print("%-60s AR %.2f ms / %d images  (%.3f ms/position)  codes: %s" % (setting or "(defaults)", best, B, best / (H * W), same),
              flush=True)
    except Exception as ex:  # noqa: BLE001
        print("%-60s FAILED: %s" % (setting, ex), flush=True)
    for k, _ in kv:
        os.environ.pop(k, None)
