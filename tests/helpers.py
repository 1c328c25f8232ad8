"""shared test helpers: build product models with synthetic weights (oracle/synth.py) on a device"""
import torch

from oracle import synth
from oracle.zoo import AR_ZOO, VAE_ZOO, vae_ddconfig
from rqvae.models import create_model
from rqvae.utils.config import Config, augment_arch_defaults


def ar_config(name):
    E, nh, nb, nhl, V, bs, vc, cl = AR_ZOO[name]
    cfg = Config(type="rq-transformer", vocab_size=V, block_size=list(bs), vocab_size_cond=vc, block_size_cond=cl,
                 embed_dim=E, input_embed_dim=256, shared_tok_emb=True, shared_cls_emb=True, input_emb_vqvae=True,
                 head_emb_vqvae=True, cumsum_depth_ctx=True,
                 body=dict(n_layer=nb, block=dict(n_head=nh)), head=dict(n_layer=nhl, block=dict(n_head=nh)))
    return augment_arch_defaults(cfg)


def vae_config(name):
    kw = VAE_ZOO[name]
    cs = kw.get("code_shape", (8, 8, 4))
    cfg = Config(type="rq-vae", hparams=dict(bottleneck_type="rq", embed_dim=256, n_embed=kw["K"],
                                             latent_shape=[cs[0], cs[1], 256], code_shape=list(cs), shared_codebook=True,
                                             decay=0.99, restart_unused_codes=True, loss_type="mse", latent_loss_weight=0.25),
                 ddconfig=vae_ddconfig(**kw))
    return augment_arch_defaults(cfg)


_AR_CACHE = {}


def build_ar(name, layouts, seed, device="cuda"):
    """synthetic-weight transformer; the large shapes are memoised per test session (regenerating 3.9 B seeded weights on the CPU
    takes minutes) -- tests set model.precision / the engine options they need explicitly"""
    key = (name, seed, str(device))
    if key in _AR_CACHE:
        model, sd = _AR_CACHE[key]
        model._invalidate_native()
        model.precision = None
        model.noise_budget_bytes = 256 << 20
        return model, sd
    with torch.device("meta"):
        model, _ = create_model(ar_config(name))
    sd = synth.synth_state_dict(layouts["ar/" + name], seed)
    model = model.to_empty(device=device)
    model.load_state_dict({k: v.to(device) for k, v in sd.items()})
    model = model.eval()
    if AR_ZOO[name][0] >= 1024:
        _AR_CACHE[key] = (model, sd)
    return model, sd


def build_vae(name, layouts, seed, device="cuda"):
    with torch.device("meta"):
        model, _ = create_model(vae_config(name))
    sd = synth.synth_state_dict(layouts["vae/" + name], seed)
    model = model.to_empty(device=device)
    model.load_state_dict({k: v.to(device) for k, v in sd.items()})
    return model.eval(), sd


class CodebookAux:
    """stand-in for an RQ-VAE that only carries the shared codebook (what RQTransformer.sample needs from model_aux)"""

    class _Q:
        shared_codebook = True

        def __init__(self, table):
            self._t = table

        def _shared_table(self):
            return self._t

    def __init__(self, table):
        self.quantizer = CodebookAux._Q(table)


def noise_tensor(seed, n_tok, B, V, device="cuda"):
    return torch.stack([synth.exp_noise(seed, t, B, V) for t in range(n_tok)]).to(device)
