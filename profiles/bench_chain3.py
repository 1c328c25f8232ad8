"""chain2 over a grid of (ctas, threads): what makes the 64 x 384 stage cost 2.5 us where the 144 x 192 one costs 1.3?
usage: python profiles/bench_chain3.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae import _native as N  # noqa: E402

L = N.lib()
ws = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")


def run(variant, words, ctas, threads, smem, n=400, reps=10):
    us = C.c_float()
    N.check(L.rqb200_dbg_chain2(variant, words, n, ctas, threads, smem, reps, ws.data_ptr(), ws.numel(), C.byref(us)), "dbg_chain2")
    return us.value


print("us / stage:  empty | R 4KB | R 64KB | W 4KB | R+W 16KB")
for ctas in (32, 64, 74, 128, 148, 296, 592, 1184):
    for threads in (64, 128, 192, 256, 384, 512):
        if ctas * threads > 148 * 2048:
            continue
        r = [run(0, 1024, ctas, threads, 0), run(1, 1024, ctas, threads, 0), run(1, 16384, ctas, threads, 0), run(2, 1024, ctas, threads, 0),
             run(3, 4096, ctas, threads, 0)]
        print("ctas %4d threads %3d : %s" % (ctas, threads, "  ".join("%5.2f" % x for x in r)), flush=True)
