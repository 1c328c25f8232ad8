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
