"""where does the time of one dependent stage go?  (csrc/dbg_chain.cu chain2)   usage: python profiles/bench_chain2.py
variant: R = read data another CTA of the previous kernel wrote, W = write, Rs = read clean lines nobody writes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae import _native as N  # noqa: E402

L = N.lib()
ws = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
NAMES = {0: "empty", 1: "R", 2: "W", 3: "R+W", 5: "Rs", 7: "Rs+W"}


def run(variant, words, ctas, threads, smem, n=400, reps=10):
    us = C.c_float()
    N.check(L.rqb200_dbg_chain2(variant, words, n, ctas, threads, smem, reps, ws.data_ptr(), ws.numel(), C.byref(us)), "dbg_chain2")
    return us.value


for ctas, threads, smem in ((144, 192, 200 << 10), (64, 384, 0), (148, 256, 0)):
    for kb in (1, 4, 16, 64):
        row = []
        for variant in (0, 1, 2, 3, 5, 7):
            row.append("%s %5.2f" % (NAMES[variant], run(variant, kb * 256, ctas, threads, smem)))
        print("ctas %3d threads %3d smem %3d KB  %2d KB/CTA : %s  us/stage" % (ctas, threads, smem >> 10, kb, "  ".join(row)), flush=True)
