"""(NOT RUNNABLE ANY MORE: needs the removed rqb200_set_conv_sm_limit entry point and the high-priority graph capture; kept as the record of
what overlap_r2.txt measured.)  Pipelining experiment: decode of batch i on a second stream (persistent conv kernels confined to `dec_sms` SMs) while batch i+1 is
sampled on the first stream (GEMM grids shrunk to what is left).   usage: python profiles/exp_overlap.py [dec_sms ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
from rqvae import _native as N  # noqa: E402

B, K = 64, 6
name = "in1400m"
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
ar, vae, dd = bench.build_models(name, dev, "fast")
E, nh, nb, nhl, V, bs, vc, cl = bench.MODELS[name][:8]
part = torch.zeros(B, *bs, dtype=torch.long, device=dev)
cond = torch.randint(0, vc, (B, cl), device=dev)
L = N.lib()


def sample():
    return ar.sample(part, model_aux=vae, cond=cond, top_k=1024, amp=True)


def sequential(k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        px = vae.decode_code(sample())
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


def pipelined(k):
    sa, sb = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prev = None
    for _ in range(k):
        with torch.cuda.stream(sa):
            codes = sample()
            ev = torch.cuda.Event()
            ev.record(sa)
        if prev is not None:
            with torch.cuda.stream(sb):
                sb.wait_event(prev[1])
                prev[0].record_stream(sb)
                px = vae.decode_code(prev[0])
        prev = (codes, ev)
    with torch.cuda.stream(sb):
        sb.wait_event(prev[1])
        prev[0].record_stream(sb)
        px = vae.decode_code(prev[0])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


def set_splits(q, p, f1, f2):
    for k, v in (("QKV", q), ("PROJ", p), ("FC1", f1), ("FC2", f2)):
        if v:
            os.environ["RQB200_SPLIT_" + k] = str(v)
        else:
            os.environ.pop("RQB200_SPLIT_" + k, None)
    ar._invalidate_native()


for it in range(2):
    sequential(1)
print("sequential (148 SMs each): %.1f ms / step = %.1f images/s" % ((ms := sequential(K)), B / ms * 1e3), flush=True)
for spec in (sys.argv[1:] or ["40:3,9,2,9", "48:2,8,2,8", "56:2,7,2,7", "148:0,0,0,0"]):
    dec_sms, splits = spec.split(":")
    set_splits(*[int(x) for x in splits.split(",")])
    L.rqb200_set_conv_sm_limit(int(dec_sms))
    try:
        pipelined(2)
        ms = pipelined(K)
        L.rqb200_set_conv_sm_limit(0)
        only_ar = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            sample()
        torch.cuda.synchronize()
        only_ar = (time.perf_counter() - t0) / 3 * 1e3
        L.rqb200_set_conv_sm_limit(int(dec_sms))
        c = sample()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            vae.decode_code(c)
        torch.cuda.synchronize()
        only_dec = (time.perf_counter() - t0) / 3 * 1e3
        print("decode on %s SMs, splits %s: pipelined %.1f ms / step = %.1f images/s   (alone: AR %.1f ms, decode %.1f ms)"
              % (dec_sms, splits, ms, B / ms * 1e3, only_ar, only_dec), flush=True)
    except Exception as ex:  # noqa: BLE001
        print("decode on %s SMs, splits %s: FAILED %s" % (dec_sms, splits, ex), flush=True)
    L.rqb200_set_conv_sm_limit(0)
