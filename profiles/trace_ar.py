"""Stage trace of the fast AR tier (RQB200_TRACE=1): where the time of one spatial position goes.
Every kernel's CTA 0 stamps %globaltimer at entry, when its upstream dependency resolves (griddepcontrol.wait returns), at
a kernel-specific midpoint (GEMM: accumulator complete) and when it is done; the stamps of the LAST replay of each captured
graph are read back.  Prints, per kernel name, the average of: dep->done (the dependent part of the stage), done(prev)->dep
(hand-over between consecutive kernels), entry->dep (how early the kernel was resident = prefetch window).
usage: python profiles/trace_ar.py [model] [B]   -> gpurun_out/trace_ar_<model>.csv"""
import os
import sys

os.environ.setdefault("RQB200_TRACE", "1")      # "2": GEMM stamp 0 = prefetched weights landed (instead of entry)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rq-vae-transformer_b200"))
import torch  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "in1400m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
ar, vae, dd = bench.build_models(name, dev, "fast", tiny_vae=True)
E, nh, nb, nhl, V, bs, vc, cl = bench.MODELS[name][:8]
part = torch.zeros(B, *bs, dtype=torch.long, device=dev)
cond = torch.randint(0, max(vc, 1), (B, cl), device=dev)
for it in range(2):
    torch.manual_seed(1)
    codes = ar.sample(part, model_aux=vae, cond=cond, top_k=min(1024, V), amp=True)
torch.cuda.synchronize()
rows = ar.native_trace()
rows.sort(key=lambda r: r[5])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = os.path.join(ROOT, "gpurun_out", "trace_ar_%s.csv" % name)
with open(out, "w") as f:
    f.write("slot,name,t_entry,t_dep,t_mid,t_done\n")
    for n, t0, t1, t2, t3, i in rows:
        f.write("%d,%s,%d,%d,%d,%d\n" % (i, n, t0, t1, t2, t3))
# per graph (slot quarter): consecutive launches
agg = {}
by_graph = {}
for r in rows:
    by_graph.setdefault(r[5] // 1024, []).append(r)
for gi, rs in sorted(by_graph.items()):
    span = (rs[-1][4] - rs[0][2]) / 1e3
    print("graph %d: %d traced launches, first dep -> last done %.1f us (%.2f us / launch)" % (gi, len(rs), span, span / len(rs)))
    for prev, cur in zip(rs[:-1], rs[1:]):
        n = cur[0]
        a = agg.setdefault((gi, n), [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += (cur[4] - cur[2]) / 1e3          # dep -> done
        a[2] += (cur[2] - prev[4]) / 1e3         # prev done -> dep resolved
        a[3] += (cur[2] - cur[1]) / 1e3          # entry -> dep (resident and waiting: prefetch window)
        a[4] += (cur[3] - cur[2]) / 1e3          # dep -> mid
print("%-3s %-12s %5s %10s %12s %12s %10s" % ("g", "kernel", "n", "dep->done", "prevdone->dep", "entry->dep", "dep->mid"))
for (gi, n), a in sorted(agg.items()):
    print("%-3d %-12s %5d %10.2f %12.2f %12.2f %10.2f" % (gi, n, a[0], a[1] / a[0], a[2] / a[0], a[3] / a[0], a[4] / a[0]))
print("wrote", out)
