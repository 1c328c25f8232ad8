"""what does one dependent stage cost?  (csrc/dbg_chain.cu)   usage: python profiles/bench_chain.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae import _native as N  # noqa: E402

L = N.lib()
ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
NAMES = {0: "PDL chain, empty kernels", 1: "PDL chain, 16 KB L2 round trip / CTA", 2: "persistent, grid barrier",
         3: "persistent, point-to-point flags", 4: "PDL chain, 16 KB, loads before stores"}
MODES = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4]


def run(mode, ctas, threads, smem, fan, n=400, reps=10):
    us = C.c_float()
    N.check(L.rqb200_dbg_chain(mode, n, ctas, threads, smem, fan, reps, ws.data_ptr(), ws.numel(), C.byref(us)), "dbg_chain")
    return us.value


for mode in MODES:
    for ctas, threads, smem in ((144, 192, 0), (144, 192, 100 << 10), (144, 192, 200 << 10), (64, 384, 0), (288, 128, 0), (444, 128, 0)):
        for fan in ((1, 8, 32) if mode else (1,)):
            try:
                us = run(mode, ctas, threads, smem, fan)
                print("mode %d  %-40s ctas %3d  threads %3d  smem %3d KB  fan %2d : %6.2f us / stage" %
                      (mode, NAMES[mode], ctas, threads, smem >> 10, fan, us), flush=True)
            except N.NativeError as ex:
                print("mode %d ctas %d smem %d: %s" % (mode, ctas, smem >> 10, str(ex)[-60:]), flush=True)
