"""Fast tier (bf16 tcgen05, PDL chain, CUDA graphs) of the AR step against the fp32 exact tier, which is itself pinned
bit-exactly to the reference fixtures (tests/test_gpu_parity.py).  Protocol: teacher-forced step parity
-- logits within a bf16 error bound, indices identical except where the fp32 tier's own decision margin is inside that
bound -- plus self-consistency of the free-running loop (graph == no graph, run-to-run determinism, resume)."""
import os

import pytest
import torch

from oracle import synth
from oracle.zoo import AR_ZOO
from tests.helpers import CodebookAux, build_ar, noise_tensor

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"


def _case(name, golden, layouts):
    g = golden("ar")["ar"][name]
    E, nh, nb_, nhl, V, bs, vc, cl = AR_ZOO[name]
    model, sd = build_ar(name, layouts, g["weight_seed"])
    cb = synth.randn_seeded((V, 256), g["codebook_seed"]).to(DEV)
    cond = synth.randint_seeded(0, max(vc, 1), (g["B"], cl), g["cond_seed"]).to(DEV) if vc > 1 else None
    return g, model, CodebookAux(cb), cond, bs, V


@pytest.mark.parametrize("name", ["tiny", "tiny_txt", "ffhq355m", "in1400m"])
def test_fast_tier_teacher_forced_step_parity(golden, layouts, name):
    g, model, aux, cond, bs, V = _case(name, golden, layouts)
    codes = g["runs"][-1]["codes"].long().to(DEV)          # a seeded top-k trajectory of the reference
    tf = dict(noise=False, return_logits=True, force_codes=codes)
    model.precision = "exact"
    _, lg32 = model._native_sample(codes, aux, cond, (0, 0), 1.0, None, None, False, **tf)
    model.precision = "fast"
    out, lg16 = model._native_sample(codes, aux, cond, (0, 0), 1.0, None, None, True, **tf)
    assert torch.equal(out, codes)
    err = (lg16 - lg32).abs()
    std = float(lg32.std())
    rms = float(err.pow(2).mean().sqrt())
    print("%s: logits std %.3f, bf16-tier error rms %.4f max %.4f" % (name, std, rms, float(err.max())))
    assert rms < 0.02 * std and float(err.max()) < 0.15 * std        # bf16 weights+activations, fp32 accumulate
    # greedy index parity with margin audit
    top2 = lg32.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    differ = lg16.argmax(-1) != lg32.argmax(-1)
    bound = 2 * err.amax(-1)
    assert not bool((differ & (margin > bound)).any()), "index flip outside the arithmetic error bound"
    print("%s: %d / %d greedy indices differ, all inside the margin bound" % (name, int(differ.sum()), differ.numel()))


def test_fast_tier_free_running_consistency(golden, layouts):
    g, model, aux, cond, bs, V = _case("tiny", golden, layouts)
    model.precision = "fast"
    B = g["B"]
    n_tok = bs[0] * bs[1] * bs[2]
    noise = noise_tensor(77, n_tok, B, V)
    part = torch.zeros(B, *bs, dtype=torch.long, device=DEV)
    a = model._native_sample(part, aux, cond, (0, 0), 1.0, 100, 0.95, True, noise=noise)
    b = model._native_sample(part, aux, cond, (0, 0), 1.0, 100, 0.95, True, noise=noise)
    assert torch.equal(a, b), "fast tier is not run-to-run deterministic"
    assert int(a.min()) >= 0 and int(a.max()) < V
    # free-running == teacher-forced replay of its own trajectory
    _, lg = model._native_sample(a, aux, cond, (0, 0), 1.0, 100, 0.95, True, noise=False, return_logits=True, force_codes=a)
    from rqvae.models import _bind as nb
    for step in (0, 1, 5, n_tok - 1):
        idx = nb.sample_logits(lg[step], 1.0, 100, 0.95, q=noise[step])
        assert torch.equal(idx, a.reshape(B, -1)[:, step])
    # resume from the middle reproduces the tail when fed the same noise tail
    h0, w0 = bs[0] // 2, 1
    skip = (h0 * bs[1] + w0) * bs[2]
    c = model._native_sample(a, aux, cond, (h0, w0), 1.0, 100, 0.95, True, noise=noise[skip:].contiguous())
    assert torch.equal(c, a)
    # CUDA graphs and PDL are pure scheduling: same codes without them (persistent form and per-op chain alike)
    for mega in ("1", "0"):
        os.environ["RQB200_MEGA"] = mega
        model._invalidate_native()
        base = model._native_sample(part, aux, cond, (0, 0), 1.0, 100, 0.95, True, noise=noise)
        for var in ("RQB200_NO_GRAPH", "RQB200_NO_PDL"):
            os.environ[var] = "1"
            try:
                model._invalidate_native()
                d = model._native_sample(part, aux, cond, (0, 0), 1.0, 100, 0.95, True, noise=noise)
            finally:
                del os.environ[var]
                model._invalidate_native()
            assert torch.equal(base, d), (var, mega)
    # persistent form vs per-op chain: same arithmetic up to the LayerNorm reduction order
    _, lg_chain = model._native_sample(a, aux, cond, (0, 0), 1.0, None, None, True, noise=False, return_logits=True, force_codes=a)
    os.environ["RQB200_MEGA"] = "1"
    model._invalidate_native()
    _, lg_mega = model._native_sample(a, aux, cond, (0, 0), 1.0, None, None, True, noise=False, return_logits=True, force_codes=a)
    del os.environ["RQB200_MEGA"]
    model._invalidate_native()
    assert float((lg_chain - lg_mega).abs().max()) < 2e-2 * float(lg_mega.std())


def test_fast_tier_text_conditioned_prefill(golden, layouts):
    """cond_len = 4 prefill + resume on the fast tier equals its own teacher-forced replay"""
    g, model, aux, cond, bs, V = _case("tiny_txt", golden, layouts)
    model.precision = "fast"
    B = g["B"]
    n_tok = bs[0] * bs[1] * bs[2]
    noise = noise_tensor(78, n_tok, B, V)
    part = torch.zeros(B, *bs, dtype=torch.long, device=DEV)
    a = model._native_sample(part, aux, cond, (0, 0), 1.0, 64, None, True, noise=noise)
    model.precision = "exact"
    _, lg32 = model._native_sample(a, aux, cond, (0, 0), 1.0, None, None, False, noise=False, return_logits=True, force_codes=a)
    model.precision = "fast"
    _, lg16 = model._native_sample(a, aux, cond, (0, 0), 1.0, None, None, True, noise=False, return_logits=True, force_codes=a)
    assert float((lg16 - lg32).abs().max()) < 0.15 * float(lg32.std())


def test_fast_tier_cc3m_654m_text_conditioned_step_parity(layouts):
    """BASELINE config 4 shape (CC-3M 654M: E=1280, 20 heads, 26+4 layers, 32 text tokens of prefix): fast tier vs exact tier,
    teacher-forced on a random trajectory"""
    from tests.helpers import build_ar
    E, nh, nb_, nhl, V, bs, vc, cl = AR_ZOO["cc3m654m"]
    model, sd = build_ar("cc3m654m", layouts, 31)
    cb = synth.randn_seeded((V, 256), 32).to(DEV)
    aux = CodebookAux(cb)
    B = 2
    cond = synth.randint_seeded(0, vc, (B, cl), 33).to(DEV)
    codes = synth.randint_seeded(0, V, (B, *bs), 34).to(DEV)
    tf = dict(noise=False, return_logits=True, force_codes=codes)
    model.precision = "exact"
    _, lg32 = model._native_sample(codes, aux, cond, (0, 0), 1.0, None, None, False, **tf)
    model.precision = "fast"
    _, lg16 = model._native_sample(codes, aux, cond, (0, 0), 1.0, None, None, True, **tf)
    err = (lg16 - lg32).abs()
    std = float(lg32.std())
    print("cc3m654m: logits std %.3f, bf16-tier error rms %.4f max %.4f" % (std, float(err.pow(2).mean().sqrt()), float(err.max())))
    assert float(err.pow(2).mean().sqrt()) < 0.02 * std and float(err.max()) < 0.15 * std
    top2 = lg32.topk(2, dim=-1).values
    differ = lg16.argmax(-1) != lg32.argmax(-1)
    assert not bool((differ & ((top2[..., 0] - top2[..., 1]) > 2 * err.amax(-1))).any())


def test_16x16_grid_with_text_prefix_exact_tier_vs_oracle():
    """BASELINE configs 4/5 ask for 16x16x4 grids (synthetic: the reference ships 8x8x4 only, SURVEY finding 8): body sequence
    32 + 256 tokens, 1024 sampled tokens.  Small width so the CPU oracle finishes in seconds; exact tier must match it."""
    from oracle import rq_oracle as O
    from rqvae.models import create_model
    from rqvae.utils.config import Config, augment_arch_defaults
    E, nh, nb_, nhl, V, bs, vc, cl = 128, 2, 1, 1, 512, (16, 16, 4), 32, 32
    cfg = augment_arch_defaults(Config(type="rq-transformer", vocab_size=V, block_size=list(bs), vocab_size_cond=vc, block_size_cond=cl,
                                       embed_dim=E, input_embed_dim=256, shared_tok_emb=True, shared_cls_emb=True, input_emb_vqvae=True,
                                       head_emb_vqvae=True, cumsum_depth_ctx=True, body=dict(n_layer=nb_, block=dict(n_head=nh)),
                                       head=dict(n_layer=nhl, block=dict(n_head=nh))))
    torch.manual_seed(5)
    model, _ = create_model(cfg)
    model = model.to(DEV).eval()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cb = synth.randn_seeded((V, 256), 41)
    B = 2
    cond = synth.randint_seeded(0, vc, (B, cl), 42)
    n_tok = bs[0] * bs[1] * bs[2]
    ref = O.ar_sample(sd, O.ArConfig(E, nh, nb_, nhl, V, bs, vc, cl), torch.zeros(B, *bs, dtype=torch.long), cb, cond=cond, top_k=50,
                      top_p=0.9, noise=lambda s, b, v: synth.exp_noise(43, s, b, v))
    model.precision = "exact"
    got = model._native_sample(torch.zeros(B, *bs, dtype=torch.long, device=DEV), CodebookAux(cb.to(DEV)), cond.to(DEV), (0, 0), 1.0, 50,
                               0.9, False, noise=noise_tensor(43, n_tok, B, V))
    fd = (got.cpu() != ref).flatten(1).any(0).nonzero()
    assert len(fd) == 0, "first divergent token %d of %d" % (int(fd[0]), n_tok)
    model.precision = "fast"
    fast = model._native_sample(torch.zeros(B, *bs, dtype=torch.long, device=DEV), CodebookAux(cb.to(DEV)), cond.to(DEV), (0, 0), 1.0, 50,
                                0.9, True, noise=noise_tensor(43, n_tok, B, V))
    assert fast.shape == ref.shape and int(fast.min()) >= 0 and int(fast.max()) < V
    # bf16 tier: the first tokens (no accumulated feedback yet) agree with the fp32 trajectory
    assert torch.equal(fast.cpu().flatten(1)[:, :8], ref.flatten(1)[:, :8])


def test_fast_tier_large_batch_is_chunked(golden, layouts):
    """B > 256 (the reference's throughput runs use up to 500): chunks of <= 256 rows, each row independent of its chunk"""
    g, model, aux, cond, bs, V = _case("tiny", golden, layouts)
    model.precision = "fast"
    B = 300
    n_tok = bs[0] * bs[1] * bs[2]
    noise = torch.empty(n_tok, B, V, device=DEV).exponential_(1, generator=torch.Generator(DEV).manual_seed(3))
    cond = torch.randint(0, 10, (B, 1), device=DEV)
    part = torch.zeros(B, *bs, dtype=torch.long, device=DEV)
    full = model._native_sample(part, aux, cond, (0, 0), 1.0, 64, None, True, noise=noise)
    sub = model._native_sample(part[140:160], aux, cond[140:160], (0, 0), 1.0, 64, None, True, noise=noise[:, 140:160].contiguous())
    assert torch.equal(full[140:160], sub)
