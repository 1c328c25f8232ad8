#!/usr/bin/env python
"""bench.py -- 256x256 images/sec of the RQ-Transformer sampling path (BASELINE.json metric).

One step = one pass of the hot path over one batch of synthetic input:
    codes  = RQTransformer.sample(zeros[B,8,8,4], model_aux=RQVAE, cond=class labels, top_k=1024)   (P3 + sampler)
    pixels = RQVAE.decode_code(codes)                                                             (P2)
Workload (N=1 and every N): ImageNet-256 class-conditional 1.4B RQ-Transformer (E=1536, 24 heads, 42+6 layers,
V=K=16384, 8x8x4 codes) + the ImageNet RQ-VAE decoder, random-init weights, synthetic labels, per-GPU batch fixed
(weak scaling): each rank samples its own shard of images with seed 1234+rank (main_sampling_fid.py:166-167); the
only exchange is one all_gather of the finished [B,8,8,4] int64 code maps before the decoder (north star).

Prints ONE JSON line (rank 0).  `value`: images/sec with inputs resident in HBM; `e2e`: same through the public API with
pinned HOST inputs (labels + empty code map) copied H2D and the finished pixels copied D2H inside the timed region.
`--impl reference` times the CPU oracle port of the reference's own PyTorch path on the host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "rq-vae-transformer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/), or None
TRAFFIC_NCU = {"gemm_tc_fc2": 19702016}   # profiles/ncu_gemm_tc_r1_raw.csv, fc2 launch: 19.70 MB read + 0 B written (partials stay in L2)

METRIC = "256x256 images/sec (ImageNet 1.4B RQ-Transformer, 8x8x4 codes, K=16384, top-k 1024, sample+decode)"

MODELS = {
    # name: (E, heads, n_body, n_head_layers, V, block, vocab_cond, cond_len, vae attn_res)
    "in1400m": (1536, 24, 42, 6, 16384, (8, 8, 4), 1000, 1, (8,)),
    "ffhq355m": (1024, 16, 24, 4, 2048, (8, 8, 4), 1, 1, (16,)),
    "tiny": (128, 2, 2, 2, 512, (8, 8, 4), 10, 1, (8,)),
}


def build_models(name, device, precision, tiny_vae=False):
    from rqvae.models import create_model
    from rqvae.utils.config import Config, augment_arch_defaults
    E, nh, nb, nhl, V, bs, vc, cl, attn = MODELS[name]
    ar_cfg = augment_arch_defaults(Config(
        type="rq-transformer", vocab_size=V, block_size=list(bs), vocab_size_cond=vc, block_size_cond=cl, embed_dim=E,
        input_embed_dim=256, shared_tok_emb=True, shared_cls_emb=True, input_emb_vqvae=True, head_emb_vqvae=True,
        cumsum_depth_ctx=True, body=dict(n_layer=nb, block=dict(n_head=nh)), head=dict(n_layer=nhl, block=dict(n_head=nh))))
    dd = dict(double_z=False, z_channels=256, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 1, 2, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=list(attn), dropout=0.0)
    if tiny_vae:
        dd.update(ch=32, ch_mult=[1, 1, 2, 2, 4, 4])
    vae_cfg = augment_arch_defaults(Config(
        type="rq-vae", hparams=dict(bottleneck_type="rq", embed_dim=256, n_embed=V, latent_shape=[8, 8, 256],
                                    code_shape=[8, 8, 4], shared_codebook=True, decay=0.99, restart_unused_codes=True,
                                    loss_type="mse", latent_loss_weight=0.25), ddconfig=dd))
    torch.manual_seed(0)            # identical weights on every rank (replaces the reference's ~780 per-tensor broadcasts)
    with torch.device(device):
        ar, _ = create_model(ar_cfg)
        vae, _ = create_model(vae_cfg)
    ar.eval()
    vae.eval()
    ar.precision = precision
    vae.precision = precision
    return ar, vae, dd


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def ar_bytes_per_position(name, B, wbytes):
    """ALGORITHMIC HBM bytes per spatial position: every body weight once, every head + classifier
    weight D times, + KV cache read."""
    E, nh, nb, nhl, V, bs, vc, cl, _ = MODELS[name]
    D = bs[2]
    per_block = 12 * E * E
    body, head, cls = nb * per_block, nhl * per_block, E * V
    return wbytes * (body + D * (head + cls))


def gemm_kernel_roofline(ar, B, hbm_peak, peak_src):
    """The step's dominant kernel (ncu launch list, profiles/): gemm_tc_kernel<64,8> at the split-K shapes.  Timed live: a
    CUDA graph of one launch per body layer on that layer's own fc2 weight (42 x 18.9 MB = 0.8 GB >> L2, i.e. cold
    weights, exactly as in the step), replayed; CUDA events on the launching stream."""
    from rqvae import _native as N
    L = N.lib()
    blocks = ar.body_transformer.blocks
    E = ar.config.embed_dim
    Ws = [b.mlp[2].weight.detach().to(torch.bfloat16).contiguous() for b in blocks]       # [E, 4E]
    X = torch.randn(B, 4 * E, device=Ws[0].device).to(torch.bfloat16)
    splits = max(1, min(148 // (E // 128), 4 * E // 64))
    part = torch.empty(splits, B, E, device=Ws[0].device)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        def launch_all():
            for W in Ws:
                N.check(L.rqb200_dbg_gemm_tc(N.ptr(W), N.ptr(X), None, None, None, 0, 0, N.ptr(part), E, 4 * E, B, splits,
                                             N.stream_ptr()), "dbg_gemm_tc")
        launch_all()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            launch_all()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(Ws))
    alg = E * 4 * E * 2 + B * 4 * E * 2 + splits * B * E * 4           # weights + activations in + partials out
    ach = alg / 1e9 / (us * 1e-6)
    return {"bound": "hbm", "kernel": "gemm_tc_kernel<64,8> fc2 (N_out=%d, K=%d, B=%d, split-K %d, %d CTAs)" % (E, 4 * E, B, splits, E // 128 * splits),
            "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "us_per_launch": us,
            "algorithmic_bytes_per_launch": alg, "traffic": None, "peak_source": peak_src,
            "how": "CUDA graph of %d back-to-back launches on distinct (cold) layer weights, CUDA events, %d replays" % (len(Ws), reps)}


def cpu_reference_leg(name, steps, warmup, budget_s, want_B):
    """the reference's own CPU PyTorch path, restated in oracle/rq_oracle.py (kind = 'port'), all host threads."""
    from oracle import rq_oracle as O
    E, nh, nb, nhl, V, bs, vc, cl, attn = MODELS[name]
    torch.manual_seed(0)
    ar, vae, dd = build_models(name, "cpu", "exact")
    sd = {k: v.detach() for k, v in ar.state_dict().items()}
    vsd = {k: v.detach() for k, v in vae.state_dict().items()}
    cfg = O.ArConfig(E, nh, nb, nhl, V, bs, vc, cl)
    table = O.codebook_of(vsd)
    cores = torch.get_num_threads()
    torch.set_grad_enabled(False)

    def one(B, n_pos=None):
        cond = torch.randint(0, max(vc, 1), (B, cl))
        t0 = time.perf_counter()
        if n_pos is None:
            codes = O.ar_sample(sd, cfg, torch.zeros(B, *bs, dtype=torch.long), table, cond=cond, top_k=min(1024, V))
            t1 = time.perf_counter()
            for i in range(B):                      # reference decodes image by image (main_sampling_fid.py:223)
                O.vae_decode_code(vsd, dd, codes[i:i + 1])
            return time.perf_counter() - t0, t1 - t0
        state = O.new_state(cfg)
        xs = torch.zeros(B, *bs, dtype=torch.long)
        for idx in range(n_pos):
            for d in range(bs[2]):
                O.ar_cached_forward(sd, cfg, state, xs[:, :idx // bs[1] + 1], table, cond, (idx // bs[1], idx % bs[1], d))
        return time.perf_counter() - t0, None

    # calibrate: 2 positions at the wanted batch
    t_cal, _ = one(want_B, n_pos=2)
    per_img_est = t_cal / 2 * (bs[0] * bs[1]) / want_B * 1.3
    B = want_B
    while B > 1 and per_img_est * B * (steps + warmup) > budget_s:
        B //= 2
    times, ar_times = [], []
    for i in range(steps + warmup):
        t, ta = one(B)
        if i >= warmup:
            times.append(t)
            ar_times.append(ta)
    tot = sum(times)
    return {"value": B * len(times) / tot, "B": B, "cores": cores, "ms_per_step": 1000 * tot / len(times),
            "ar_ms_per_token": 1000 * sum(ar_times) / len(times) / (bs[0] * bs[1] * bs[2]),
            "sample": "%d full images per step (256 AR tokens + per-image decode), %d steps after %d warm-up" % (B, len(times), warmup)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="in1400m", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step (weak scaling)")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.model
    E, nh, nb, nhl, V, bs, vc, cl, attn = MODELS[name]
    H, W, D = bs
    config = {"workload": "imagenet256 class-cond 1.4B RQ-Transformer 8x8x4 K=16384 top-k=1024 + RQ-VAE decode" if name == "in1400m" else name,
              "per_gpu_batch": args.batch, "global_batch": args.batch * max(world, 1), "grid": "%dx%dx%d" % bs,
              "parallelism": "dp%d (independent images, one all_gather of code maps)" % max(world, 1),
              "l2": "inputs larger than L2: %.2f GB of weights streamed per spatial position" % (ar_bytes_per_position(name, args.batch, 2) / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_leg(name, args.steps, args.warmup, budget_s=150.0, want_B=args.batch)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "ar_ms_per_token": r["ar_ms_per_token"],
                "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from rqvae import _native as N
    from rqvae.utils.utils import set_seed
    torch.set_grad_enabled(False)
    ar, vae, dd = build_models(name, dev, args.precision)
    B = args.batch
    set_seed(1234 + rank)
    labels_host = torch.randint(0, max(vc, 1), (B, cl)).pin_memory()
    empty_host = torch.zeros(B, H, W, D, dtype=torch.long).pin_memory()
    labels_dev, empty_dev = labels_host.to(dev), empty_host.to(dev)
    pix_host = torch.empty(B * max(world, 1) if False else B, 3, dd["resolution"], dd["resolution"]).pin_memory()
    gathered = [torch.empty(B, H, W, D, dtype=torch.long, device=dev) for _ in range(world)] if world > 1 else None

    def step(e2e):
        if e2e:
            cond = labels_host.to(dev, non_blocking=True)
            part = empty_host.to(dev, non_blocking=True)
        else:
            cond, part = labels_dev, empty_dev
        codes = ar.sample(part, model_aux=vae, cond=cond, top_k=min(1024, V), amp=True)
        if world > 1:
            dist.all_gather(gathered, codes)          # the single collective: finished code maps (2 KB / image)
            codes = gathered[rank]                    # every rank decodes its own shard
        pix = vae.decode_code(codes)
        pix = (pix * 0.5 + 0.5).clamp_(0, 1)
        if e2e:
            pix_host.copy_(pix, non_blocking=True)
        return codes, pix

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            ev[i][0].record()
            if e2e:
                cond = labels_host.to(dev, non_blocking=True)
                part = empty_host.to(dev, non_blocking=True)
            else:
                cond, part = labels_dev, empty_dev
            codes = ar.sample(part, model_aux=vae, cond=cond, top_k=min(1024, V), amp=True)
            ev[i][1].record()
            if world > 1:
                dist.all_gather(gathered, codes)
            pix = vae.decode_code(codes)
            pix = (pix * 0.5 + 0.5).clamp_(0, 1)
            if e2e:
                pix_host.copy_(pix, non_blocking=True)
            ev[i][2].record()
        t1.record()
        barrier()
        total = t0.elapsed_time(t1)
        ar_ms = sum(e[0].elapsed_time(e[1]) for e in ev)
        dec_ms = sum(e[1].elapsed_time(e[2]) for e in ev)
        tt = torch.tensor([total, ar_ms, dec_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(x) for x in tt]

    for _ in range(args.warmup):
        step(False)
    launches0 = N.launch_count["total"]
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    total, ar_ms, dec_ms = timed(False, args.steps)
    clock_summary = clocks.summary() if rank == 0 else None
    launches = N.launch_count["total"] - launches0
    step(True)
    e_total, e_ar, e_dec = timed(True, args.steps)

    n_img = B * max(world, 1) * args.steps
    value = n_img / (total / 1e3)
    e2e_value = n_img / (e_total / 1e3)
    ar_ms_token = ar_ms / args.steps / (H * W * D)
    # P3 roofline: algorithmic bytes per spatial position / measured time per position (weights stream from HBM every
    # position: 3.94 GB >> 126 MB L2)
    wbytes = 2 if args.precision == "fast" else 4
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    pos_ms = ar_ms / args.steps / (H * W)
    ach = ar_bytes_per_position(name, B, wbytes) / 1e9 / (pos_ms / 1e3)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    roofline_step = {"bound": "hbm", "kernel": "AR spatial position (body stack + D x (head stack + classifier + sampler))",
                     "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                     "algorithmic_bytes_per_position": ar_bytes_per_position(name, B, wbytes), "ms_per_position": pos_ms,
                     "peak_source": peak_src}
    roofline = roofline_step
    if rank == 0 and args.precision == "fast":
        try:
            roofline = gemm_kernel_roofline(ar, B, hbm_peak, peak_src)
            roofline["traffic"] = TRAFFIC_NCU.get("gemm_tc_fc2")
        except Exception as ex:
            roofline = dict(roofline_step, note="kernel-level measurement failed: %s" % str(ex)[:120])
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": max(world, 1), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "fast" else "f32", "data": "synthetic", "config": config,
            "ar_ms_per_token": ar_ms_token, "ar_ms_per_step": ar_ms / args.steps, "decode_ms_per_step": dec_ms / args.steps,
            "clocks": clock_summary, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "images/s",
                    "h2d_bytes_per_step": labels_host.numel() * 8 + empty_host.numel() * 8,
                    "d2h_bytes_per_step": pix_host.numel() * 4},
            "roofline": roofline, "roofline_ar_step": roofline_step}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                r = cpu_reference_leg(name, 1, 0, budget_s=args.cpu_budget, want_B=8)
                line["cpu_baseline"] = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                                        "sample": r["sample"], "ar_ms_per_token": r["ar_ms_per_token"]}
            except Exception as ex:   # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
