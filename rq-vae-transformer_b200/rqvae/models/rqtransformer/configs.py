"""Transformer config defaults (reference: rqvae/models/rqtransformer/configs.py:21-73).

The reference declares these as OmegaConf structured dataclasses, which no longer import on Python >= 3.11
(dataclass-instance defaults, SURVEY.md finding 2).  Here they are plain nested dicts merged by
``rqvae.utils.config.Config`` -- same field names, same defaults, same ``create`` classmethod."""
from ...utils.config import Config, MISSING

ATTENTION_BLOCK_DEFAULTS = dict(embed_dim=MISSING, n_head=MISSING, mlp_bias=True, attn_bias=True, attn_pdrop=0.0,
                                resid_pdrop=0.1, gelu="v1")
ATTENTION_STACK_DEFAULTS = dict(n_layer=MISSING, block=ATTENTION_BLOCK_DEFAULTS)
RQTRANSFORMER_DEFAULTS = dict(
    type="rq-transformer", ema=None, ar_hierarchy=None, vocab_size=MISSING, block_size=MISSING, vocab_size_cond=0,
    block_size_cond=0, embed_dim=MISSING, input_embed_dim=None, use_padding_emb=False, input_emb_vqvae=False,
    head_emb_vqvae=False, scaled_head_emb_vqvae=False, cumsum_depth_ctx=False, shared_tok_emb=False, embd_pdrop=0.0,
    body=ATTENTION_STACK_DEFAULTS, head=ATTENTION_STACK_DEFAULTS, shared_cls_emb=False)


class AttentionBlockConfig(Config):
    def __init__(self, *a, **kw):
        super().__init__(ATTENTION_BLOCK_DEFAULTS)
        self.update_from(dict(*a, **kw))


class AttentionStackConfig(Config):
    def __init__(self, *a, **kw):
        super().__init__(ATTENTION_STACK_DEFAULTS)
        self.update_from(dict(*a, **kw))


class RQTransformerConfig(Config):
    def __init__(self, *a, **kw):
        super().__init__(RQTRANSFORMER_DEFAULTS)
        self.update_from(dict(*a, **kw))

    @classmethod
    def create(cls, config):
        """configs.py:68-73 -- defaults with body/head block width tied to embed_dim, overridden by ``config``"""
        out = cls(embed_dim=config["embed_dim"] if isinstance(config, dict) else config.embed_dim)
        out.body.block.embed_dim = out.embed_dim
        out.head.block.embed_dim = out.embed_dim
        out.update_from(config)
        return out
