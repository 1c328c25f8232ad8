"""does it matter WHO wrote the lines a stage reads?  chain2 with contiguous (one writer per region) vs scattered (every region
written by all CTAs) producers.   usage: python profiles/bench_chain4.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae import _native as N  # noqa: E402

L = N.lib()
ws = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")


def run(variant, words, ctas, threads, smem=0, n=400, reps=10):
    us = C.c_float()
    N.check(L.rqb200_dbg_chain2(variant, words, n, ctas, threads, smem, reps, ws.data_ptr(), ws.numel(), C.byref(us)), "dbg_chain2")
    return us.value


print("us / stage:   R+W contiguous | R+W scattered | W contiguous | W scattered")
for ctas, threads in ((148, 256), (148, 384), (144, 192), (74, 384)):
    for kb in (4, 16, 32, 64):
        w = kb * 256
        print("ctas %3d threads %3d %2d KB/CTA : %5.2f  %5.2f  %5.2f  %5.2f" %
              (ctas, threads, kb, run(3, w, ctas, threads), run(11, w, ctas, threads), run(2, w, ctas, threads), run(10, w, ctas, threads)),
              flush=True)
