"""per-SM shared-memory fill rate from L2: tensor-map boxes (rows x 128 B) vs 1-D bulk copies (csrc/dbg_tma.cu)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
from rqvae import _native as N  # noqa: E402

L = N.lib()
buf = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
for ctas in (148, 36, 1):
    for rows, depth in ((64, 8), (128, 8), (64, 1), (128, 1), (256, 4)):
        for boxes in (8, 96, 768):
            row = []
            for mode in (0, 1):
                bpc, us = C.c_float(), C.c_float()
                N.check(L.rqb200_dbg_tma_rate(mode, rows, depth, 200, boxes, N.ptr(buf), ctas, C.byref(bpc), C.byref(us)), "dbg_tma_rate")
                row.append((bpc.value, us.value))
            print("ctas %3d  box %3d rows x 128 B  depth %d  %4d distinct boxes : tensor-map %6.1f B/clk/SM (%6.2f us/round)   "
                  "1-D bulk %6.1f B/clk/SM (%6.2f us/round)" % (ctas, rows, depth, boxes, row[0][0], row[0][1], row[1][0], row[1][1]),
                  flush=True)
