"""Pins the CPU oracle (oracle/rq_oracle.py) against golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py, run in the build container).  CPU only."""
import pytest
import torch

from oracle import rq_oracle as O
from oracle import synth
from oracle.zoo import AR_ZOO, VAE_ZOO, vae_ddconfig

torch.set_grad_enabled(False)


def test_rq_quantize_matches_reference(golden):
    for c in golden("rq")["rq"]:
        cb = synth.randn_seeded((c["K"], 256), 1000 + c["seed"])
        x = synth.randn_seeded((c["B"], 8, 8, 256), 2000 + c["seed"], 0.2)
        quants, codes = O.rq_quantize(x, cb, 4)
        assert torch.equal(codes.to(torch.int32), c["codes"])
        assert torch.equal(quants[-1][:, ::4, ::4, ::16], c["agg_last_sub"])
        for q, s in zip(quants, c["agg_sum"]):
            assert abs(float(q.double().sum()) - s) < 1e-6
        assert torch.equal(O.embed_code(codes, cb)[:, ::4, ::4, ::16], c["embed_sub"])


def test_rq_ties_first_index_wins(golden):
    cb = synth.randn_seeded((64, 256), 4242)
    cb[40] = cb[5]
    cb[63] = cb[5]
    x = synth.randn_seeded((1, 8, 8, 256), 4243, 0.2)
    x[0, 0, 0] = cb[5]
    x[0, 0, 1] = cb[40] * 1.0
    _, codes = O.rq_quantize(x, cb, 4)
    assert torch.equal(codes.to(torch.int32), golden("rq")["rq_ties"]["codes"])
    assert int(codes[0, 0, 0, 0]) == 5 and int(codes[0, 0, 1, 0]) == 5
    assert not bool(((codes == 40) | (codes == 63)).any())


def test_sampler_matches_reference(golden):
    n = 0
    for c in golden("sampler")["sampler"]:
        logits = synth.randn_seeded((c["B"], c["V"]), c["seed"], 2.5 if not c.get("ties") else 1.0)
        if c.get("ties"):
            logits[:, 100:140] = 1.25
            logits[1, 7] = 30.0
        q = synth.exp_noise(c["seed"], 0, c["B"], c["V"])
        idx = O.sample_from_logits(logits, c["T"], c["k"], 1.0 if c["p"] is None else c["p"], q=q)
        assert torch.equal(idx.to(torch.int32), c["idx"]), c
        n += 1
    assert n >= 20


def _ar_setup(name, g):
    E, nh, nb, nhl, V, bs, vc, cl = AR_ZOO[name]
    cfg = O.ArConfig(E, nh, nb, nhl, V, bs, vc, cl)
    return cfg, V, bs, vc, cl


@pytest.mark.parametrize("name", ["tiny", "tiny_txt"])
def test_ar_sample_matches_reference(golden, layouts, name):
    g = golden("ar")["ar"][name]
    cfg, V, bs, vc, cl = _ar_setup(name, g)
    sd = synth.synth_state_dict(layouts["ar/" + name], g["weight_seed"])
    cb = synth.randn_seeded((V, 256), g["codebook_seed"])
    cond = synth.randint_seeded(0, max(vc, 1), (g["B"], cl), g["cond_seed"]) if vc > 1 else None
    for run in g["runs"]:
        kept = {}
        codes = O.ar_sample(sd, cfg, torch.zeros(g["B"], *bs, dtype=torch.long), cb, cond=cond,
                            noise=lambda step, B, V_, s=run["noise_seed"]: synth.exp_noise(s, step, B, V_),
                            logits_hook=lambda step, loc, lg: kept.__setitem__(step, lg.clone()), **run["setting"])
        assert torch.equal(codes.to(torch.int32), run["codes"])
        if run["logits"]:
            for step, lg in run["logits"].items():
                torch.testing.assert_close(kept[step], lg, rtol=1e-5, atol=1e-5)
    rs = g["resume"]
    codes2 = O.ar_sample(sd, cfg, g["runs"][0]["codes"].long(), cb, cond=cond, start_loc=rs["start_loc"],
                         top_k=rs["top_k"], noise=lambda step, B, V_: synth.exp_noise(rs["noise_seed"], step, B, V_))
    assert torch.equal(codes2.to(torch.int32), rs["codes"])


@pytest.mark.slow
def test_ar_355m_first_rows_match_reference(golden, layouts):
    """full-size FFHQ-355M: teacher-forced check of the stored logits + free-running greedy prefix."""
    name = "ffhq355m"
    g = golden("ar")["ar"][name]
    cfg, V, bs, vc, cl = _ar_setup(name, g)
    sd = synth.synth_state_dict(layouts["ar/" + name], g["weight_seed"])
    cb = synth.randn_seeded((V, 256), g["codebook_seed"])
    run = g["runs"][0]
    ref_codes = run["codes"].long()
    # teacher-forced: feed the reference's own codes, compare logits at the stored steps (<= step 5 to stay fast)
    state = O.new_state(cfg)
    step = 0
    for h in range(1):
        for w in range(2):
            for d in range(4):
                lg = O.ar_cached_forward(sd, cfg, state, ref_codes[:, :h + 1], cb, None, (h, w, d))
                if step in run["logits"]:
                    torch.testing.assert_close(lg, run["logits"][step], rtol=1e-5, atol=2e-5)
                assert torch.equal(lg.argmax(-1), ref_codes[:, h, w, d])          # greedy run
                step += 1


@pytest.mark.parametrize("name", ["tiny", "tiny_attn_mid"])
def test_vae_matches_reference(golden, layouts, name):
    _check_vae(golden, layouts, name)


@pytest.mark.slow
def test_vae_imagenet_decode_matches_reference(golden, layouts):
    _check_vae(golden, layouts, "imagenet", encode=False)


def _check_vae(golden, layouts, name, encode=True):
    g = golden("vae")["vae"][name]
    kw = VAE_ZOO[name]
    dd = vae_ddconfig(**kw)
    K = kw["K"]
    cs = kw.get("code_shape", (8, 8, 4))
    sd = synth.synth_state_dict(layouts["vae/" + name], g["weight_seed"])
    codes = synth.randint_seeded(0, K, (2, *cs), g["codes_seed"])
    st = g["stride"]
    pix = O.vae_decode_code(sd, dd, codes)
    torch.testing.assert_close(pix[:, :, ::st, ::st], g["pixels_sub"], rtol=1e-5, atol=1e-5)
    assert abs(float(pix.double().pow(2).sum().sqrt()) - g["pixels_l2"]) < 1e-3 * g["pixels_l2"]
    if encode:
        x = synth.randn_seeded((2, 3, dd["resolution"], dd["resolution"]), g["x_seed"])
        z_e = O.vae_encode(sd, dd, x)
        torch.testing.assert_close(z_e, g["z_e"], rtol=1e-5, atol=1e-5)
        out, codes_fwd, _ = O.vae_forward(sd, dd, x, cs[2])
        assert torch.equal(codes_fwd.to(torch.int32), g["codes_fwd"])
        torch.testing.assert_close(out[:, :, ::st, ::st], g["recon_sub"], rtol=1e-4, atol=1e-4)
