"""RQTransformer -- host-side mirror of rqvae/models/rqtransformer/transformers.py:34-369 over the native AR engine.

Boundary kept: constructor from an ``RQTransformerConfig``-shaped object, parameter names
(state_dict layout A.3), ``sample`` :294-307 (same signature; ``fast``/``cached``/``is_tqdm``/``desc`` accepted),
``cached_forward`` :191, ``init_cache`` :289, ``get_block_size``, attributes ``block_size`` / ``block_size_cond`` /
``vocab_size``.  The (h,w,d) loop, KV caches, embedding glue, classifier and ``sample_from_logits`` all run inside
``rqb200_ar_sample`` (csrc/ar_engine.cu): one native call per ``sample``, no per-token host work.

Arithmetic tiers: ``amp=False`` -> 'exact' (fp32 weights/activations, FFMA -- the tier the bit-exact-indices gate is
defined on); ``amp=True`` -> 'fast' (fp16 weights / activations / KV on tcgen05 tensor cores, fp32 accumulate -- the
reference's own amp class is fp16 autocast, transformers.py:114,206; RQB200_FAST_DTYPE=bf16 selects bf16 instead).
``self.precision`` ('exact' | 'fast') or RQB200_PRECISION overrides the ``amp`` mapping."""
import ctypes as C
import os
from collections import OrderedDict
from itertools import product

import torch
import torch.nn as nn

from ... import _native as N
from ..interfaces import Stage2Model


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("rqb200: parameter holder -- compute runs in the native engine (RQTransformer.sample)")


class _Attn(_Holder):
    def __init__(self, E, n_head, bias):
        super().__init__()
        self.key, self.query, self.value = nn.Linear(E, E, bias=bias), nn.Linear(E, E, bias=bias), nn.Linear(E, E, bias=bias)
        self.proj = nn.Linear(E, E, bias)
        self.n_head = n_head


class _Block(_Holder):
    def __init__(self, cfg):
        super().__init__()
        E = cfg.embed_dim
        assert E % cfg.n_head == 0
        self.ln1, self.ln2 = nn.LayerNorm(E), nn.LayerNorm(E)
        self.attn = _Attn(E, cfg.n_head, cfg.attn_bias)
        # indices 0 / 2 carry the weights (attentions.py:117-122); 1 / 3 are GELU / dropout in the reference
        self.mlp = nn.Sequential(nn.Linear(E, 4 * E, bias=cfg.mlp_bias), nn.Identity(), nn.Linear(4 * E, E, bias=cfg.mlp_bias),
                                 nn.Identity())
        if cfg.gelu != "v1":
            raise NotImplementedError("rqb200: only the exact-erf GELU ('v1') is implemented (all shipped configs)")


class _Stack(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.blocks = nn.ModuleList([_Block(cfg.block) for _ in range(cfg.n_layer)])


class RQTransformer(Stage2Model):
    def __init__(self, config):
        super().__init__()
        self.config = config = config.copy()
        if len(config.block_size) != 3:
            raise ValueError("incompatible block size")
        self.block_size = torch.Size(config.block_size)
        if isinstance(config.vocab_size, int):
            config.vocab_size = [config.vocab_size] * config.block_size[2]
        vs = list(config.vocab_size)
        if not (config.shared_tok_emb and config.shared_cls_emb and config.input_emb_vqvae and config.head_emb_vqvae
                and config.cumsum_depth_ctx):
            raise NotImplementedError("rqb200: only the shipped configuration family is implemented (shared_tok_emb, "
                                      "shared_cls_emb, input_emb_vqvae, head_emb_vqvae, cumsum_depth_ctx all true)")
        assert [vs[0]] * len(vs) == vs
        self.vocab_size = vs
        E = config.embed_dim
        self.vocab_size_cond = max(config.vocab_size_cond, 1)
        self.block_size_cond = max(config.block_size_cond, 1)
        assert not (self.block_size_cond > 1 and self.vocab_size_cond == 1)
        self.cond_emb = nn.Embedding(self.vocab_size_cond, E)
        self.tok_emb = None
        self.input_mlp = nn.Linear(config.input_embed_dim, E)
        self.head_mlp = nn.Linear(config.input_embed_dim, E)
        self.pos_emb_cond = nn.Parameter(torch.zeros(1, self.block_size_cond, E))
        self.pos_emb_hw = nn.Parameter(torch.zeros(1, self.block_size[0] * self.block_size[1], E))
        self.pos_emb_d = nn.Parameter(torch.zeros(1, self.block_size[2], E))
        for p in (self.pos_emb_cond, self.pos_emb_hw, self.pos_emb_d):
            p.data.normal_(mean=0.0, std=0.02)
        self.body_transformer = _Stack(config.body)
        self.head_transformer = _Stack(config.head)
        self.classifier = nn.Sequential(OrderedDict([("layer_norm", nn.LayerNorm(E)), ("linear", nn.Linear(E, vs[0]))]))
        if config.block_size_cond > 1:
            self.cond_classifier = nn.Sequential(OrderedDict([("layer_norm", nn.LayerNorm(E)),
                                                              ("linear", nn.Linear(E, config.vocab_size_cond))]))
        self.precision = None
        self.noise_budget_bytes = 256 << 20      # bound on the Exp(1) noise buffer sample() draws per span of positions
        self._eng = {}
        self._eng_fp = None
        self._cache = None
        self.last_launches = 0

    # ------------------------------------------------------------------ native engine plumbing
    def _invalidate_native(self):
        for e in self._eng.values():
            N.lib().rqb200_ar_destroy(e["handle"])
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

    def _mode(self, amp):
        p = self.precision or N.default_precision()
        if p == "auto":
            p = "fast" if amp else "exact"
        return N.MODE_FAST if p == "fast" else N.MODE_EXACT

    @staticmethod
    def _codebook_of(model_aux, depth):
        """the [K,C] table behind model_aux.get_code_emb_with_depth (transformers.py:109-111)"""
        q = getattr(model_aux, "quantizer", None)
        if q is not None and hasattr(q, "_shared_table"):
            return q._shared_table()
        if q is not None and getattr(q, "shared_codebook", False):
            return q.codebooks[0].weight[:-1]
        raise NotImplementedError("rqb200: model_aux must be an RQ-VAE with a shared codebook")

    def _engine(self, codebook, mode, slot=0):
        dev = self.pos_emb_hw.device
        fp = (N.param_fingerprint(self), codebook._version)
        if fp != self._eng_fp:               # weights changed behind the module's own hooks (wrapper load, in-place write)
            self._invalidate_native()
            self._eng_fp = fp
        key = (str(dev), mode, codebook.data_ptr(), slot)
        if key not in self._eng and slot > 0:
            # engines of one model share the packed weights of slot 0; each slot owns its workspace, KV cache and graphs
            base = self._engine(codebook, mode, 0)
            eng = {"handle": base["make"](), "keep": base["keep"], "ws": None, "make": base["make"]}
            self._eng[key] = eng
            return eng
        if key in self._eng:
            return self._eng[key]
        N.require_cuda(self.pos_emb_hw, codebook)
        L = N.lib()
        wdt = N.fast_dtype() if mode == N.MODE_FAST else torch.float32
        opts = N.ar_engine_options() if mode == N.MODE_FAST else {"flags": 0, "splits": [0, 0, 0, 0]}
        keep = []

        def f32(t):
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def wt(t):
            t = t.detach().to(wdt).contiguous()
            keep.append(t)
            return t.data_ptr()

        def blocks(stack):
            arr = (N.BlockWeights * len(stack.blocks))()
            for i, b in enumerate(stack.blocks):
                a = b.attn
                wqkv = torch.cat([a.query.weight, a.key.weight, a.value.weight], 0)
                bqkv = torch.cat([a.query.bias, a.key.bias, a.value.bias], 0)
                arr[i].wqkv, arr[i].bqkv = wt(wqkv), f32(bqkv)
                arr[i].w1, arr[i].b1 = wt(b.mlp[0].weight), f32(b.mlp[0].bias)
                arr[i].wproj, arr[i].bproj = wt(a.proj.weight), f32(a.proj.bias)
                arr[i].w2, arr[i].b2 = wt(b.mlp[2].weight), f32(b.mlp[2].bias)
                arr[i].ln1_w, arr[i].ln1_b = f32(b.ln1.weight), f32(b.ln1.bias)
                arr[i].ln2_w, arr[i].ln2_b = f32(b.ln2.weight), f32(b.ln2.bias)
            return arr

        cfg = N.ArConfig()
        c = self.config
        cfg.embed_dim, cfg.n_head = c.embed_dim, c.body.block.n_head
        cfg.n_body, cfg.n_head_layers = len(self.body_transformer.blocks), len(self.head_transformer.blocks)
        cfg.vocab, cfg.H, cfg.W, cfg.D = self.vocab_size[0], self.block_size[0], self.block_size[1], self.block_size[2]
        cfg.vocab_cond, cfg.cond_len = self.vocab_size_cond, self.block_size_cond
        cfg.code_dim, cfg.codebook_size = codebook.shape[1], codebook.shape[0]
        cfg.mode, cfg.weight_dtype = mode, N._DT[wdt]
        cfg.flags = opts["flags"]
        cfg.split_qkv, cfg.split_proj, cfg.split_fc1, cfg.split_fc2 = opts["splits"]
        if c.head.block.n_head != c.body.block.n_head:
            raise NotImplementedError("rqb200: body and head stacks must share n_head")
        w = N.ArWeights()
        w.pos_emb_cond, w.pos_emb_hw, w.pos_emb_d = f32(self.pos_emb_cond), f32(self.pos_emb_hw), f32(self.pos_emb_d)
        w.cond_emb = f32(self.cond_emb.weight)
        w.w_in, w.b_in = wt(self.input_mlp.weight), f32(self.input_mlp.bias)
        w.w_head, w.b_head = wt(self.head_mlp.weight), f32(self.head_mlp.bias)
        w.w_cls, w.b_cls = wt(self.classifier.linear.weight), f32(self.classifier.linear.bias)
        w.cls_ln_w, w.cls_ln_b = f32(self.classifier.layer_norm.weight), f32(self.classifier.layer_norm.bias)
        w.codebook = f32(codebook)
        if hasattr(self, "cond_classifier") and mode == N.MODE_FAST:
            pad = -self.vocab_size_cond % 128            # classifier rows padded with zeros up to the 128-feature tcgen05 tile
            pw = torch.nn.functional.pad(self.cond_classifier.linear.weight.detach(), (0, 0, 0, pad))
            pb = torch.nn.functional.pad(self.cond_classifier.linear.bias.detach(), (0, pad))
            w.w_ccls, w.b_ccls = wt(pw), f32(pb)
            w.ccls_ln_w, w.ccls_ln_b = f32(self.cond_classifier.layer_norm.weight), f32(self.cond_classifier.layer_norm.bias)
        body, head = blocks(self.body_transformer), blocks(self.head_transformer)
        w.body, w.head = C.cast(body, C.POINTER(N.BlockWeights)), C.cast(head, C.POINTER(N.BlockWeights))
        keep.extend([body, head, cfg, w])

        def make():
            hnd = L.rqb200_ar_create(C.byref(cfg), C.byref(w))
            if not hnd:
                raise N.NativeError("rqb200_ar_create: " + L.rqb200_last_error().decode())
            return hnd

        handle = make()
        eng = {"handle": handle, "keep": keep, "ws": None, "make": make, "dtype": wdt}
        self._eng[key] = eng
        return eng

    # ------------------------------------------------------------------ reference surface
    def init_cache(self):
        """transformers.py:289-292 -- KV state lives inside one native call; nothing persists between calls"""
        self._cache = {"spatial_ctx_hw": None}

    def _lists(self, top_k, top_p):
        D = self.block_size[2]
        V = self.vocab_size
        if top_k is None:
            ks = [V[i] for i in range(D)]
        elif isinstance(top_k, int):
            ks = [min(top_k, V[i]) for i in range(D)]
        elif len(top_k) == 1:
            ks = [min(top_k[0], V[i]) for i in range(D)]
        else:
            ks = [min(top_k[i], V[i]) for i in range(D)]
        if top_p is None:
            ps = [1.0] * D
        elif isinstance(top_p, float):
            ps = [min(top_p, 1.0)] * D
        elif len(top_p) == 1:
            ps = [min(top_p[0], 1.0)] * D
        else:
            ps = [min(top_p[i], 1.0) for i in range(D)]
        return ks, ps

    @torch.no_grad()
    def _native_sample(self, partial, model_aux, cond, start_loc, temperature, top_k, top_p, amp, noise=None,
                       return_logits=False, force_codes=None):
        H, W, D = self.block_size
        B = partial.shape[0]
        dev = self.pos_emb_hw.device
        N.require_cuda(partial, cond, self.pos_emb_hw)
        ks, ps = self._lists(top_k, top_p)
        codebook = self._codebook_of(model_aux, D)
        mode = self._mode(amp)
        partial = partial.to(torch.int64).contiguous()
        cond_t = None if cond is None else cond.reshape(B, self.block_size_cond).to(torch.int64).contiguous()
        idx0 = start_loc[0] * W + start_loc[1]
        n_tok = max(H * W - idx0, 0) * D
        V = self.vocab_size[0]
        HWD = H * W * D
        with torch.cuda.device(dev):
            draw = noise is None            # draw the Exp(1) noise here, exactly as torch.multinomial would (utils.py:114)
            if noise is False:
                noise = None
            logits = torch.empty(n_tok, B, V, dtype=torch.float32, device=dev) if return_logits else None
            out = torch.empty_like(partial)
            kk = (C.c_int32 * D)(*[int(k) for k in ks])
            pp = (C.c_float * D)(*[float(p) for p in ps])
            fc = None if force_codes is None else force_codes.to(torch.int64).contiguous()
            bounds = [(0, B)]
            if mode == N.MODE_FAST and B > 256:
                # the tcgen05 tier takes at most 256 batch rows per call (UMMA N <= 256): run equal chunks back to back
                if return_logits:
                    raise N.NativeError("rqb200: return_logits with B > 256 is not supported on the fast tier")
                n_chunks = -(-B // 256)
                bounds = [(i * B // n_chunks, (i + 1) * B // n_chunks) for i in range(n_chunks)]
            # position spans: when the noise is drawn here it is drawn span by span into one bounded buffer (noise_budget_bytes)
            # instead of one [n_tok,B,V] tensor (1 GB at 8x8x4, B=64, V=16384); every batch chunk keeps its own engine slot
            # (workspace + KV state) so that all chunks can resume on the next span
            n_pos = H * W - idx0
            per_pos = max(1, D * B * V * 4)
            span = max(1, min(n_pos, int(self.noise_budget_bytes) // per_pos)) if draw else max(n_pos, 1)
            if draw and n_tok > 0:
                noise = torch.empty(min(span, n_pos) * D, B, V, dtype=torch.float32, device=dev)
            st = torch.cuda.current_stream(dev)
            launches = 0
            engines = []
            for slot, (lo, hi) in enumerate(bounds):
                eng = self._engine(codebook, mode, slot)
                need = N.lib().rqb200_ar_workspace_bytes(eng["handle"], hi - lo)
                if eng["ws"] is None or eng["ws"].numel() < need:
                    eng["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
                engines.append(eng)

            def off(t, lo, row_elems, esize, extra=0):
                return C.c_void_p(t.data_ptr() + (lo * row_elems + extra) * esize) if t is not None else C.c_void_p(0)

            for p0 in range(idx0, H * W, span):
                p1 = min(p0 + span, H * W)
                if draw:
                    # one exponential_ per token, in (h,w,d) order: the draws torch.multinomial would make
                    for t in range((p1 - p0) * D):
                        noise[t].exponential_(1)
                tok0 = 0 if draw else (p0 - idx0) * D                  # first token of this span inside `noise`
                for eng, (lo, hi) in zip(engines, bounds):
                    N.check(N.lib().rqb200_ar_sample_span(
                        eng["handle"], off(partial, lo, HWD, 8), off(cond_t, lo, self.block_size_cond, 8), hi - lo, p0, p1,
                        int(p0 > idx0), float(temperature), kk, pp, off(noise, lo, V, 4, tok0 * B * V),
                        0 if noise is None else B * V, off(logits, lo, V, 4, (p0 - idx0) * D * B * V), off(fc, lo, HWD, 8),
                        off(out, lo, HWD, 8), N.ptr(eng["ws"]), eng["ws"].numel(), C.c_void_p(st.cuda_stream)), "ar_sample")
                    launches += N.lib().rqb200_ar_last_launches(eng["handle"])
            if n_tok == 0:
                out.copy_(partial)
        self.last_launches = launches
        N.launch_count["total"] += launches
        return (out, logits) if return_logits else out

    def native_trace(self):
        """diagnostics (RQB200_TRACE=1, fast tier): [(name, t_entry, t_dependency_resolved, t_mid, t_done)] in ns for every launch
        slot of the last replay of each captured graph -- where the time of one AR position goes"""
        rows = []
        for key, eng in self._eng.items():
            cap = 4096
            buf = (C.c_longlong * (4 * cap))()
            names = C.create_string_buffer(cap * 16)
            n = N.lib().rqb200_ar_trace(eng["handle"], buf, cap, names, len(names))
            nm = names.value.decode().split("\n")
            for i in range(max(n, 0)):
                if buf[4 * i]:
                    rows.append((nm[i] if i < len(nm) else "?", buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3], i))
        return rows

    @torch.no_grad()
    def sample(self, partial_sample, model_aux=None, cond=None, start_loc=(0, 0), temperature=1.0, top_k=None, top_p=None,
               amp=False, cached=True, is_tqdm=False, desc="Sampling", fast=True):
        """transformers.py:294-369.  Returns LongTensor [B,H,W,D]; ``partial_sample`` is not modified."""
        assert self.block_size == partial_sample.shape[1:]
        self.init_cache()
        out = self._native_sample(partial_sample, model_aux, cond, start_loc, temperature, top_k, top_p, amp)
        self.init_cache()
        return out

    @torch.no_grad()
    def cached_forward(self, xs, model_aux=None, cond=None, amp=False, sample_loc=(0, 0, 0)):
        """transformers.py:190-287 -- logits [B,V] for one (h,w,d).  Stateless re-evaluation: the prefix in ``xs`` is
        teacher-forced through the native loop and the requested step's logits are returned."""
        h, w, d = sample_loc
        H, W, D = self.block_size
        B = xs.shape[0]
        full = torch.zeros(B, H, W, D, dtype=torch.int64, device=xs.device)
        full[:, :xs.shape[1]] = xs
        _, logits = self._native_sample(full, model_aux, cond, (0, 0), 1.0, None, None, amp, noise=False,
                                        return_logits=True, force_codes=full)
        return logits[(h * W + w) * D + d]

    def forward(self, xs, model_aux=None, cond=None, amp=False):
        """transformers.py:113-188 -- teacher-forced logits [B,H,W,D,V]; with cond_len > 1 also the cond logits
        [B,cond_len-1,vocab_cond] (the reference's return convention :185-188).
        Fast tier (amp=True): all positions at once -- the body over B*(cond_len+H*W-1) rows, the head over B*H*W*D rows, as
        large-M tcgen05 GEMMs + causal attention (rqb200_ar_forward).  Exact tier: the sequential teacher-forced replay."""
        B, H, W, D = xs.shape
        if self._mode(amp) == N.MODE_FAST:
            return self._native_forward(xs, model_aux, cond)
        if self.block_size_cond > 1:
            raise NotImplementedError("rqb200: cond_logits (cond_len > 1) are produced by the fast tier only (amp=True)")
        _, logits = self._native_sample(xs, model_aux, cond, (0, 0), 1.0, None, None, amp, noise=False,
                                        return_logits=True, force_codes=xs)
        return logits.reshape(H, W, D, B, -1).permute(3, 0, 1, 2, 4).contiguous()

    @torch.no_grad()
    def _native_forward(self, xs, model_aux, cond):
        H, W, D = self.block_size
        B = xs.shape[0]
        dev = self.pos_emb_hw.device
        N.require_cuda(xs, cond, self.pos_emb_hw)
        codebook = self._codebook_of(model_aux, D)
        eng = self._engine(codebook, N.MODE_FAST)
        xs = xs.to(torch.int64).contiguous()
        cl, V = self.block_size_cond, self.vocab_size[0]
        cond_t = None if cond is None else cond.reshape(B, cl).to(torch.int64).contiguous()
        want_cond = cl > 1 and hasattr(self, "cond_classifier")
        with torch.cuda.device(dev):
            need = N.lib().rqb200_ar_forward_workspace_bytes(eng["handle"], B)
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            logits = torch.empty(D, H * W, B, V, dtype=torch.float32, device=dev)
            vcp = -(-self.vocab_size_cond // 128) * 128
            cond_logits = torch.empty(cl - 1, B, vcp, dtype=torch.float32, device=dev) if want_cond else None
            N.check(N.lib().rqb200_ar_forward(eng["handle"], N.ptr(xs), N.ptr(cond_t), B, N.ptr(logits), N.ptr(cond_logits), N.ptr(ws),
                                              ws.numel(), N.stream_ptr(dev)), "ar_forward")
            self.last_launches = N.lib().rqb200_ar_last_launches(eng["handle"])
            N.launch_count["total"] += self.last_launches
        out = logits.permute(2, 1, 0, 3).reshape(B, H, W, D, V).contiguous()
        if want_cond:
            return out, cond_logits[..., :self.vocab_size_cond].permute(1, 0, 2).contiguous()
        return out

    def compute_loss(self, logits, targets, use_soft_target=False):
        return torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), targets.reshape(-1))
