"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's sampling hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and there only as the *checker* (or the timed CPU baseline), never as the
product path.  The product (``rq-vae-transformer_b200/``) never imports it and has no CPU fallback.

What it restates (reference = kakaobrain/rq-vae-transformer @ 341395e, paths relative to its root):

  P1  residual quantisation search .......... rqvae/models/rqvae/quantizations.py:43-69, 237-271, 297-334
  P2  conv encoder / decoder ................ rqvae/models/rqvae/{layers.py:16-182, modules.py:73-98,171-202, rqvae.py:80-109}
  P3  cached AR step + sampling loop ........ rqvae/models/rqtransformer/{transformers.py:190-369, attentions.py:60-165}
      sampler ............................... rqvae/utils/utils.py:60-123

Arithmetic: the reference has no native code; all of its arithmetic is PyTorch (third-party, not vendored in
``/root/reference``; the authors tested torch 1.9, this image ships torch 2.11).  The restatement therefore
uses the same ``torch`` CPU primitives (``addmm``, ``F.linear``, ``F.conv2d``, ``F.group_norm``, ``bmm``, ``topk``,
``sort``, ``cumsum``) in the same order on plain weight dictionaries (``state_dict`` key layout of SURVEY.md A.3),
fp32 throughout -- exactly what the reference executes on CPU, where ``autocast`` is a no-op.  Functions are
stateless; KV caches are explicit Python lists.

Pinning: the reference ships no tests or golden vectors.  This oracle is pinned against the
reference itself, imported in the build container through ``oracle/ref_loader.py``: ``oracle/gen_golden.py``
writes the fixtures under ``tests/golden/`` from the *reference classes*, and ``tests/test_oracle_golden.py``
checks this restatement against those fixtures (and, when ``/root/reference`` is present, directly against the
live reference).

RNG injection: ``torch.multinomial(probs, 1)`` == ``argmax(probs / q)`` with ``q = empty_like(probs).exponential_(1)``
(SURVEY.md finding 7).  ``sample_from_logits`` accepts the noise tensor ``q`` explicitly so that the CUDA engine and
the oracle can be fed the identical per-token noise.
"""
import math
from itertools import product

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------ P1

def vq_distances(x, codebook):
    """quantizations.py:43-62 -- ||x||^2 + ||e||^2 - 2 x.e via one addmm, fp32.  ``codebook`` = weight[:-1]."""
    cb_t = codebook.t()
    flat = x.reshape(-1, codebook.shape[1])
    xn = flat.pow(2.0).sum(dim=1, keepdim=True)
    en = cb_t.pow(2.0).sum(dim=0, keepdim=True)
    d = torch.addmm(xn + en, flat, cb_t, alpha=-2.0)
    return d.reshape(*x.shape[:-1], -1)


def rq_quantize(x, codebook, depth):
    """quantizations.py:237-271 (shared codebook).  x [B,h,w,C] -> (list of D cumulative quants, codes [B,h,w,D] int64)."""
    residual = x.detach().clone()
    agg = torch.zeros_like(x)
    quants, codes = [], []
    for _ in range(depth):
        idx = vq_distances(residual, codebook).argmin(dim=-1)          # :64-69, first index wins ties
        q = F.embedding(idx, codebook)                                 # :144-146
        residual.sub_(q)                                               # :264
        agg.add_(q)                                                    # :265
        quants.append(agg.clone())
        codes.append(idx.unsqueeze(-1))
    return quants, torch.cat(codes, dim=-1)


def rq_soft_codes(x, codebook, depth, temp=1.0):
    """quantizations.py:371-399 (stochastic=False): per-depth softmax(-distances / temp) [..., D, K] and the argmin codes."""
    residual = x.detach().clone()
    softs, codes = [], []
    for _ in range(depth):
        d = vq_distances(residual, codebook)
        softs.append(F.softmax(-d / temp, dim=-1).unsqueeze(-2))
        idx = d.argmin(dim=-1)
        residual.sub_(F.embedding(idx, codebook))
        codes.append(idx.unsqueeze(-1))
    return torch.cat(softs, dim=-2), torch.cat(codes, dim=-1)


def embed_code(codes, codebook):
    """quantizations.py:297-311 -- sum over depth of codebook rows (cat then sum(-2)); rH=rW=1 so no reshape."""
    parts = [F.embedding(c, codebook) for c in torch.chunk(codes, codes.shape[-1], dim=-1)]
    return torch.cat(parts, dim=-2).sum(-2)


def embed_code_with_depth(codes, codebook):
    """quantizations.py:313-334 -- per-depth embeddings [..., D, C] without the sum."""
    parts = [F.embedding(c, codebook) for c in torch.chunk(codes, codes.shape[-1], dim=-1)]
    return torch.cat(parts, dim=-2)


# ------------------------------------------------------------------------------------------------ sampler

def top_k_logits(logits, k):
    """utils.py:60-64 -- keeps every value >= the k-th largest (ties kept)."""
    v, _ = torch.topk(logits, k)
    out = logits.clone()
    out[out < v[:, [-1]]] = -float("inf")
    return out


def top_p_probs(probs, p):
    """utils.py:67-79 -- sort desc, cumsum, drop where the *previous* cumulative mass >= p, renormalise."""
    sp, si = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(sp, dim=-1)
    rm = cum >= p
    rm[..., 1:] = rm[..., :-1].clone()
    rm[..., 0] = 0
    rm = rm.scatter(-1, si, rm)
    probs = probs.masked_fill(rm, 0.0)
    return probs / torch.sum(probs, dim=-1, keepdim=True)


def sample_from_logits(logits, temperature=1.0, top_k=None, top_p=None, q=None):
    """utils.py:82-123.  ``q``: optional Exp(1) noise [B,V]; None -> draw it from the global generator exactly
    as ``torch.multinomial`` would (one exponential_ call of shape [B,V])."""
    logits = logits.to(dtype=torch.float32) / temperature
    if top_k is not None:
        logits = top_k_logits(logits, top_k)
    nan = torch.isnan(logits)
    if nan.any():
        logits[nan] = -float("inf")
    probs = F.softmax(logits, dim=-1)
    if top_p is not None:
        probs = top_p_probs(probs, top_p)
    if q is None:
        q = torch.empty_like(probs).exponential_(1)
    return torch.argmax(probs / q, dim=-1)


# ------------------------------------------------------------------------------------------------ P3

def _linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _layer_norm(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def attention(sd, pre, x, n_head, past_kv=None, caching=False):
    """attentions.py:60-104 -- causal multi-head self attention; optional KV cache [(B*nh,T,hs) x2]."""
    B, T, C = x.shape
    hs = C // n_head
    xt = x.transpose(0, 1).contiguous()
    k = _linear(sd, pre + ".key", xt).view(T, B * n_head, hs).transpose(0, 1)
    q = _linear(sd, pre + ".query", xt).view(T, B * n_head, hs).transpose(0, 1)
    v = _linear(sd, pre + ".value", xt).view(T, B * n_head, hs).transpose(0, 1)
    t_past = 0
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=-2)
        v = torch.cat([past_kv[1], v], dim=-2)
        t_past = past_kv[0].shape[1]
    present = torch.stack([k, v]) if caching else None
    att = torch.bmm(q, k.transpose(-2, -1) * (1.0 / math.sqrt(hs)))
    tot = t_past + T
    mask = torch.tril(torch.ones(tot, tot, dtype=torch.bool)).view(1, tot, tot)
    att = att.masked_fill(~mask[:, t_past:tot, :tot], float("-inf"))
    att = F.softmax(att, dim=-1)
    y = torch.bmm(att, v).transpose(0, 1).contiguous().view(T, B, C)
    y = _linear(sd, pre + ".proj", y).transpose(0, 1).contiguous()
    return y, present


def block(sd, pre, x, n_head, cache=None):
    """attentions.py:125-142 -- pre-LN block.  ``cache`` is a one-element list holding past_kv (or None)."""
    if cache is None:
        a, _ = attention(sd, pre + ".attn", _layer_norm(sd, pre + ".ln1", x), n_head)
    else:
        a, cache[0] = attention(sd, pre + ".attn", _layer_norm(sd, pre + ".ln1", x), n_head, cache[0], True)
    x = x + a
    h = _linear(sd, pre + ".mlp.0", _layer_norm(sd, pre + ".ln2", x))
    h = _linear(sd, pre + ".mlp.2", F.gelu(h))
    return x + h


def stack(sd, pre, x, n_layer, n_head, caches=None):
    for l in range(n_layer):
        x = block(sd, "%s.blocks.%d" % (pre, l), x, n_head, None if caches is None else caches[l])
    return x


class ArConfig:
    """plain description of an RQ-Transformer (the fields of configs.py:38-66 the sampling path reads)."""

    def __init__(self, embed_dim, n_head, n_body, n_head_layers, vocab_size, block_size=(8, 8, 4),
                 vocab_size_cond=1, block_size_cond=1, input_embed_dim=256):
        self.E, self.nh, self.n_body, self.n_headl = embed_dim, n_head, n_body, n_head_layers
        self.V, self.block_size = vocab_size, tuple(block_size)
        self.vocab_cond, self.cond_len = max(vocab_size_cond, 1), max(block_size_cond, 1)
        self.C = input_embed_dim


def ar_cached_forward(sd, cfg, state, xs, codebook, cond, loc):
    """transformers.py:190-287 -- one AR token.  ``state`` = {'ctx':None|tensor, 'body':[[kv]..], 'head':[[kv]..]}.
    Recomputes the embeddings of the whole prefix every call, exactly like the reference."""
    h, w, d = loc
    B, H, W, D = xs.shape
    idx = h * W + w
    xs = xs.clone().reshape(B, -1, D)[:, :idx + 1]
    if cond is None:
        cond = torch.zeros(B, cfg.cond_len, dtype=torch.long)
    else:
        cond = cond.reshape(B, cfg.cond_len)
    seq_len, cond_len = xs.shape[1], cond.shape[1]
    if d == 0:
        emb = _linear(sd, "input_mlp", embed_code_with_depth(xs, codebook))                    # :219-220
        c_emb = F.embedding(cond, sd["cond_emb.weight"]) + sd["pos_emb_cond"][:, :cond_len, :]  # :224
        emb = emb.sum(dim=-2) + sd["pos_emb_hw"][:, :seq_len, :]                                # :225
        lat = torch.cat([c_emb, emb[:, :-1, :]], dim=1)[:, :cond_len + idx, :]                  # :226-235
        if state["ctx"] is None:
            out = stack(sd, "body_transformer", lat, cfg.n_body, cfg.nh, state["body"])        # prefill :238
            ctx = out[:, -1, :].unsqueeze(1)
        else:
            ctx = stack(sd, "body_transformer", lat[:, -1, :].unsqueeze(1), cfg.n_body, cfg.nh, state["body"])
        state["ctx"] = ctx
    ctx = state["ctx"]
    dctx = _linear(sd, "head_mlp", torch.cumsum(embed_code_with_depth(xs, codebook), dim=-2))   # :250-255
    dctx = dctx[:, idx, :]
    full = torch.cat([ctx.view(B, 1, -1), dctx[:, :-1, :]], dim=-2) + sd["pos_emb_d"][:, :D, :]  # :260-267
    tok = full[:, d, :].unsqueeze(1)
    if d == 0:
        state["head"] = [[None] for _ in range(cfg.n_headl)]                                     # :271-272
    out = stack(sd, "head_transformer", tok, cfg.n_headl, cfg.nh, state["head"])
    logits = _linear(sd, "classifier.linear", _layer_norm(sd, "classifier.layer_norm", out))    # :278-285
    return logits.reshape(B, -1)


def _per_depth(val, default, D, cap=None):
    if val is None:
        lst = [default] * D
    elif isinstance(val, (int, float)):
        lst = [val] * D
    elif len(val) == 1:
        lst = [val[0]] * D
    else:
        lst = list(val)[:D]
    return [min(v, cap) if cap is not None else v for v in lst]


def new_state(cfg):
    return {"ctx": None, "body": [[None] for _ in range(cfg.n_body)], "head": [[None] for _ in range(cfg.n_headl)]}


def ar_sample(sd, cfg, partial_sample, codebook, cond=None, start_loc=(0, 0), temperature=1.0, top_k=None,
              top_p=None, noise=None, logits_hook=None):
    """transformers.py:294-369.  ``noise``: None (draw from the global generator, one [B,V] exponential_ per token,
    as torch.multinomial does), the string 'greedy' is not special -- pass top_k=1 for greedy -- or a callable
    ``noise(step, B, V) -> q`` / an indexable of per-token [B,V] tensors.  ``logits_hook(step, loc, logits)``."""
    H, W, D = cfg.block_size
    assert tuple(partial_sample.shape[1:]) == (H, W, D)
    ks = _per_depth(top_k, cfg.V, D, cfg.V)
    ps = _per_depth(top_p, 1.0, D, 1.0)
    xs = partial_sample.clone()
    state = new_state(cfg)
    step = 0
    for (h, w, d) in product(range(H), range(W), range(D)):
        if (h, w) < (start_loc[0], start_loc[1]):
            continue
        logits = ar_cached_forward(sd, cfg, state, xs[:, :h + 1], codebook, cond, (h, w, d))
        if logits_hook is not None:
            logits_hook(step, (h, w, d), logits)
        if noise is None:
            q = None
        elif callable(noise):
            q = noise(step, logits.shape[0], logits.shape[1])
        else:
            q = noise[step]
        xs[:, h, w, d] = sample_from_logits(logits, temperature, ks[d], ps[d], q=q)
        step += 1
    return xs


def ar_forward(sd, cfg, xs, codebook, cond=None, with_cond_logits=False):
    """transformers.py:113-188 -- teacher-forced logits [B,H,W,D,V]; with_cond_logits (cond_len > 1): also the cond_classifier
    logits of latents[:, :cond_len-1] (:153-156), the reference's two-value return convention (:185-186)."""
    B, H, W, D = xs.shape
    xs = xs.reshape(B, H * W, D)
    cond = torch.zeros(B, cfg.cond_len, dtype=torch.long) if cond is None else cond.reshape(B, cfg.cond_len)
    L, cl = xs.shape[1], cond.shape[1]
    emb = _linear(sd, "input_mlp", embed_code_with_depth(xs, codebook))
    c_emb = F.embedding(cond, sd["cond_emb.weight"]) + sd["pos_emb_cond"][:, :cl, :]
    emb = emb.sum(dim=-2) + sd["pos_emb_hw"][:, :L, :]
    lat = stack(sd, "body_transformer", torch.cat([c_emb, emb[:, :-1, :]], dim=1), cfg.n_body, cfg.nh)
    sp = lat[:, cl - 1:]
    dctx = _linear(sd, "head_mlp", torch.cumsum(embed_code_with_depth(xs, codebook), dim=-2))
    full = torch.cat([sp.view(B, L, 1, -1), dctx[:, :, :-1, :]], dim=-2).reshape(B * L, D, -1) + sd["pos_emb_d"][:, :D, :]
    out = stack(sd, "head_transformer", full, cfg.n_headl, cfg.nh).reshape(B, H, W, D, -1)
    logits = _linear(sd, "classifier.linear", _layer_norm(sd, "classifier.layer_norm", out))
    if with_cond_logits and cl > 1:
        return logits, _linear(sd, "cond_classifier.linear", _layer_norm(sd, "cond_classifier.layer_norm", lat[:, :cl - 1]))
    return logits


# ------------------------------------------------------------------------------------------------ P2

def _gn(sd, name, x):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], 1e-6)      # layers.py:16-17


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _silu(x):
    return F.silu(x)                                                                  # layers.py:11-13 (nonlinearity)


def resnet_block(sd, pre, x):
    """layers.py:100-120 (temb is None, dropout 0)."""
    h = _conv(sd, pre + ".conv1", _silu(_gn(sd, pre + ".norm1", x)), padding=1)
    h = _conv(sd, pre + ".conv2", _silu(_gn(sd, pre + ".norm2", h)), padding=1)
    if (pre + ".nin_shortcut.weight") in sd:
        x = _conv(sd, pre + ".nin_shortcut", x)
    return x + h


def attn_block(sd, pre, x):
    """layers.py:158-182 -- single-head attention over h*w tokens, scale c^-0.5, softmax over keys."""
    h_ = _gn(sd, pre + ".norm", x)
    q, k, v = _conv(sd, pre + ".q", h_), _conv(sd, pre + ".k", h_), _conv(sd, pre + ".v", h_)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, pre + ".proj_out", h_)


def decoder_forward(sd, dd, z, pre="decoder"):
    """modules.py:171-202.  ``dd`` = ddconfig dict; z [B,z_channels,h,w] NCHW."""
    nres, nblk = len(dd["ch_mult"]), dd["num_res_blocks"]
    res = dd["resolution"] // 2 ** (nres - 1)
    h = _conv(sd, pre + ".conv_in", z, padding=1)
    h = resnet_block(sd, pre + ".mid.block_1", h)
    h = attn_block(sd, pre + ".mid.attn_1", h)
    h = resnet_block(sd, pre + ".mid.block_2", h)
    for lvl in reversed(range(nres)):
        for b in range(nblk + 1):
            h = resnet_block(sd, "%s.up.%d.block.%d" % (pre, lvl, b), h)
            if res in dd["attn_resolutions"]:
                h = attn_block(sd, "%s.up.%d.attn.%d" % (pre, lvl, b), h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")                   # layers.py:31-35
            h = _conv(sd, "%s.up.%d.upsample.conv" % (pre, lvl), h, padding=1)
            res *= 2
    return _conv(sd, pre + ".conv_out", _silu(_gn(sd, pre + ".norm_out", h)), padding=1)


def encoder_forward(sd, dd, x, pre="encoder"):
    """modules.py:73-98.  x [B,3,R,R] NCHW -> [B,z_channels,r,r]."""
    nres, nblk = len(dd["ch_mult"]), dd["num_res_blocks"]
    res = dd["resolution"]
    h = _conv(sd, pre + ".conv_in", x, padding=1)
    for lvl in range(nres):
        for b in range(nblk):
            h = resnet_block(sd, "%s.down.%d.block.%d" % (pre, lvl, b), h)
            if res in dd["attn_resolutions"]:
                h = attn_block(sd, "%s.down.%d.attn.%d" % (pre, lvl, b), h)
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)                      # layers.py:50-54
            h = _conv(sd, "%s.down.%d.downsample.conv" % (pre, lvl), h, stride=2)
            res //= 2
    h = resnet_block(sd, pre + ".mid.block_1", h)
    h = attn_block(sd, pre + ".mid.attn_1", h)
    h = resnet_block(sd, pre + ".mid.block_2", h)
    return _conv(sd, pre + ".conv_out", _silu(_gn(sd, pre + ".norm_out", h)), padding=1)


def codebook_of(sd):
    """shared codebook without the zero padding row (quantizations.py:28,45)."""
    return sd["quantizer.codebooks.0.weight"][:-1]


def vae_encode(sd, dd, x):
    """rqvae.py:80-83 -> z_e [B,h,w,C] NHWC contiguous."""
    return _conv(sd, "quant_conv", encoder_forward(sd, dd, x)).permute(0, 2, 3, 1).contiguous()


def vae_decode(sd, dd, z_q):
    """rqvae.py:85-89.  z_q [B,h,w,C] NHWC -> pixels [B,3,R,R]."""
    return decoder_forward(sd, dd, _conv(sd, "post_quant_conv", z_q.permute(0, 3, 1, 2).contiguous()))


def vae_decode_code(sd, dd, codes):
    """rqvae.py:105-109."""
    return vae_decode(sd, dd, embed_code(codes, codebook_of(sd)))


def vae_forward(sd, dd, xs, depth):
    """rqvae.py:74-78 (inference values: straight-through output equals quant_list[-1] numerically up to fp32
    rounding of x + (q - x); we return exactly that expression)."""
    z_e = vae_encode(sd, dd, xs)
    quants, codes = rq_quantize(z_e, codebook_of(sd), depth)
    z_q = z_e + (quants[-1] - z_e)                                                    # quantizations.py:279
    return vae_decode(sd, dd, z_q), codes, z_e


