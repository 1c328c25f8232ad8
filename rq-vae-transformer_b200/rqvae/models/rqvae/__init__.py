from .rqvae import RQVAE


def get_rqvae(config):
    """reference: rqvae/models/rqvae/__init__.py:17-22"""
    return RQVAE(**config.hparams, ddconfig=config.ddconfig, checkpointing=getattr(config, "checkpointing", False))
