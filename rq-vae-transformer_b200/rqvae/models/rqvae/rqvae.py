"""RQVAE -- host-side mirror of rqvae/models/rqvae/rqvae.py:26-168 over the native conv engine.

Boundary kept: ``encode`` :80, ``decode`` :85, ``forward`` :74, ``get_codes`` :91, ``decode_code`` :105,
``get_code_emb_with_depth`` :146, ``decode_partial_code`` :150, ``get_recon_imgs`` :111, attribute ``code_shape`` and the
state_dict key layout.  ``precision``: 'exact' = fp32 FFMA kernels, 'fast' = fp16-operand / fp32-accumulate tcgen05
implicit GEMM (the reference's own GPU decode runs cuDNN with TF32 allowed -- same 10-bit mantissa class)."""
import ctypes as C
import os

import torch
from torch import nn
from torch.nn import functional as F

from ... import _native as N
from ..interfaces import Stage1Model
from .modules import Decoder, Encoder, ResnetBlock
from .quantizations import RQBottleneck


class RQVAE(Stage1Model):
    def __init__(self, *, embed_dim=64, n_embed=512, decay=0.99, loss_type="mse", latent_loss_weight=0.25,
                 bottleneck_type="rq", ddconfig=None, checkpointing=False, **kwargs):
        super().__init__()
        assert loss_type in ("mse", "l1")
        ddconfig = dict(ddconfig)
        self.ddconfig = ddconfig
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        if bottleneck_type != "rq":
            raise ValueError("invalid 'bottleneck_type' (must be 'rq')")
        self.quantizer = RQBottleneck(latent_shape=kwargs["latent_shape"], code_shape=kwargs["code_shape"], n_embed=n_embed,
                                      decay=decay, shared_codebook=kwargs["shared_codebook"],
                                      restart_unused_codes=kwargs["restart_unused_codes"])
        self.code_shape = kwargs["code_shape"]
        self.embed_dim = embed_dim
        self.quant_conv = nn.Conv2d(ddconfig["z_channels"], embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.loss_type, self.latent_loss_weight = loss_type, latent_loss_weight
        self.precision = None            # None -> _native.default_precision() ('auto' == exact until told otherwise)
        self.split_fp16 = True           # fast tier: 3-product split-fp16 convs (keeps 60 chained convs within 1e-3)
        self._eng = {}                   # (device, mode) -> dict(handle, tensors, ws)
        self._eng_fp = None              # parameter fingerprint the cached engines were built from
        self.last_launches = 0

    # ------------------------------------------------------------------ native engine plumbing
    def _invalidate_native(self):
        for e in self._eng.values():
            N.lib().rqb200_vae_destroy(e["handle"])
        self._eng = {}

    def _apply(self, fn, *a, **k):
        self._invalidate_native()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate_native()
        return super().load_state_dict(*a, **k)

    def __del__(self):
        try:
            self._invalidate_native()
        except Exception:
            pass

    def _mode(self):
        p = self.precision or N.default_precision()
        return N.MODE_FAST if p == "fast" else N.MODE_EXACT

    def _engine(self, device):
        mode = self._mode()
        fp = N.param_fingerprint(self)
        if fp != self._eng_fp:               # weights changed behind the module's own hooks (wrapper load, in-place write)
            self._invalidate_native()
            self._eng_fp = fp
        key = (str(device), mode, self.split_fp16)
        if key in self._eng:
            return self._eng[key]
        L = N.lib()
        dd = self.ddconfig
        cfg = N.VaeConfig()
        cfg.ch, cfg.n_levels, cfg.num_res_blocks = dd["ch"], len(dd["ch_mult"]), dd["num_res_blocks"]
        for i, m in enumerate(dd["ch_mult"]):
            cfg.ch_mult[i] = m
        cfg.n_attn_res = len(dd["attn_resolutions"])
        for i, r in enumerate(dd["attn_resolutions"]):
            cfg.attn_resolutions[i] = r
        cfg.resolution, cfg.z_channels, cfg.embed_dim = dd["resolution"], dd["z_channels"], self.embed_dim
        cfg.in_channels, cfg.out_ch = dd["in_channels"], dd["out_ch"]
        cfg.codebook_size, cfg.depth, cfg.mode = self.quantizer.n_embed[0], self.code_shape[-1], mode
        if os.environ.get("RQB200_GN_FUSE", "1") == "0":      # diagnostics: stand-alone GroupNorm statistics pass
            cfg.mode |= 0x100
        handle = L.rqb200_vae_create(C.byref(cfg))
        if not handle:
            raise N.NativeError("rqb200_vae_create: " + L.rqb200_last_error().decode())
        wdt = torch.float16 if mode == N.MODE_FAST else torch.float32
        # fast tier: the encoder runs on the tcgen05 conv path as well (every conv but the Cin = 3 conv_in); RQB200_ENC_FAST=0 keeps
        # the encoder on the fp32 kernels
        enc_fast = mode == N.MODE_FAST and os.environ.get("RQB200_ENC_FAST", "1") == "1"
        keep = {}

        def reg(name, t):
            t = t.detach().contiguous()
            keep[name] = t
            N.check(L.rqb200_vae_set_tensor(handle, name.encode(), N.ptr(t), N.dtype_code(t), t.numel()), "vae_set_tensor")

        def reg_conv(name, w_oihw):
            """conv weight OIHW -> OHWI in the engine's dtype; fast tier: fp16 hi + lo halves (split-fp16 products)"""
            w = w_oihw.detach().permute(0, 2, 3, 1).contiguous().float()
            fast16 = name.startswith(("decoder.", "post_quant_conv"))
            if enc_fast and name.startswith(("encoder.", "quant_conv")) and not name.startswith("encoder.conv_in"):
                fast16 = True
            if mode == N.MODE_FAST and fast16:
                hi = w.to(torch.float16)
                reg(name, hi)
                if self.split_fp16:
                    reg(name + "_lo", (w - hi.float()).to(torch.float16))
            else:
                reg(name, w)          # exact tier, and the encoder in every tier (fp32 FFMA kernels)

        sd = {k: v for k, v in self.state_dict().items()}
        for k, v in sd.items():
            if not (k.startswith("encoder.") or k.startswith("decoder.") or k.startswith("quant_conv") or
                    k.startswith("post_quant_conv")):
                continue
            N.require_cuda(v)
            if v.dim() == 4:                                   # conv weight OIHW -> OHWI in the engine's weight dtype
                if k.endswith((".q.weight", ".k.weight", ".v.weight")):
                    continue
                reg_conv(k, v)
            elif k.endswith((".q.bias", ".k.bias", ".v.bias")):
                continue
            else:
                reg(k, v.float())
        for k in [k for k in sd if k.endswith(".q.weight")]:   # fused q|k|v 1x1 conv (layers.py:161-163)
            base = k[:-len(".q.weight")]
            w = torch.cat([sd[base + ".q.weight"], sd[base + ".k.weight"], sd[base + ".v.weight"]], 0)
            b = torch.cat([sd[base + ".q.bias"], sd[base + ".k.bias"], sd[base + ".v.bias"]], 0)
            reg_conv(base + ".qkv.weight", w)
            reg(base + ".qkv.bias", b.float())
        reg("codebook", self.quantizer._shared_table().float())
        N.check(L.rqb200_vae_finalize(handle), "vae_finalize")
        eng = {"handle": handle, "keep": keep, "ws": {}}
        self._eng[key] = eng
        return eng

    def _ws(self, eng, B, device):
        need = N.lib().rqb200_vae_workspace_bytes(eng["handle"], B)
        ws = eng["ws"].get("buf")
        if ws is None or ws.numel() < need:
            eng["ws"]["buf"] = ws = torch.empty(need, dtype=torch.uint8, device=device)
        return ws, need

    def _run(self, fn_name, x, out_shape):
        N.require_cuda(x)
        eng = self._engine(x.device)
        B = x.shape[0]
        out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws, need = self._ws(eng, B, x.device)
            fn = getattr(N.lib(), fn_name)
            N.check(fn(eng["handle"], N.ptr(x), B, N.ptr(out), N.ptr(ws), ws.numel(), N.stream_ptr(x.device)), fn_name)
        self.last_launches = N.lib().rqb200_vae_last_launches(eng["handle"])
        N.launch_count["total"] += self.last_launches
        return out

    # ------------------------------------------------------------------ reference surface
    def _latent_hw(self):
        dd = self.ddconfig
        r = dd["resolution"] // 2 ** (len(dd["ch_mult"]) - 1)
        return r

    @torch.no_grad()
    def encode(self, x):
        """rqvae.py:80-83: [B,3,R,R] -> z_e [B,h,w,embed_dim] NHWC contiguous"""
        x = x.float().contiguous()
        r = self._latent_hw()
        return self._run("rqb200_vae_encode", x, (x.shape[0], r, r, self.embed_dim))

    @torch.no_grad()
    def decode(self, z_q):
        """rqvae.py:85-89: z_q [B,h,w,embed_dim] NHWC -> [B,out_ch,R,R]"""
        z_q = z_q.float().contiguous()
        dd = self.ddconfig
        return self._run("rqb200_vae_decode", z_q, (z_q.shape[0], dd["out_ch"], dd["resolution"], dd["resolution"]))

    @torch.no_grad()
    def forward(self, xs):
        """rqvae.py:74-78 (inference; gradients are out of scope)"""
        z_e = self.encode(xs)
        z_q, quant_loss, code = self.quantizer(z_e)
        return self.decode(z_q), quant_loss, code

    @torch.no_grad()
    def get_codes(self, xs):
        z_e = self.encode(xs)
        _, codes = self.quantizer.quantize(self.quantizer.to_code_shape(z_e))
        return codes

    @torch.no_grad()
    def decode_code(self, code):
        """rqvae.py:105-109 -- embed_code + decode in one native call"""
        assert code.shape[1:] == torch.Size(self.code_shape)
        code = code.to(torch.int64).contiguous()
        dd = self.ddconfig
        if tuple(self.quantizer.shape_divisor[:2]) != (1, 1):
            return self.decode(self.quantizer.embed_code(code))
        return self._run("rqb200_vae_decode_code", code, (code.shape[0], dd["out_ch"], dd["resolution"], dd["resolution"]))

    def get_recon_imgs(self, xs_real, xs_recon):
        return xs_real * 0.5 + 0.5, torch.clamp(xs_recon * 0.5 + 0.5, 0, 1)

    def compute_loss(self, out, quant_loss, code, xs=None, valid=False):
        """rqvae.py:119-141 (torch glue; evaluation only)"""
        loss_recon = F.mse_loss(out, xs) if self.loss_type == "mse" else F.l1_loss(out, xs)
        loss_latent = quant_loss
        if valid:
            loss_recon = loss_recon * xs.shape[0] * xs.shape[1]
            loss_latent = loss_latent * xs.shape[0]
        return {"loss_total": loss_recon + self.latent_loss_weight * loss_latent, "loss_recon": loss_recon,
                "loss_latent": loss_latent, "codes": [code]}

    def get_last_layer(self):
        return self.decoder.conv_out.weight

    @torch.no_grad()
    def get_code_emb_with_depth(self, code):
        return self.quantizer.embed_code_with_depth(code)

    @torch.no_grad()
    def decode_partial_code(self, code, code_idx, decode_type="select"):
        return self.decode(self.quantizer.embed_partial_code(code, code_idx, decode_type))

    @torch.no_grad()
    def forward_partial_code(self, xs, code_idx, decode_type="select"):
        return self.decode_partial_code(self.get_codes(xs), code_idx, decode_type)

    @torch.no_grad()
    def get_soft_codes(self, xs, temp=1.0, stochastic=False):
        """rqvae.py:97-103"""
        z_e = self.encode(xs)
        return self.quantizer.get_soft_codes(z_e, temp=temp, stochastic=stochastic)
