from .configs import RQTransformerConfig
from .transformers import RQTransformer


def get_rqtransformer(config):
    """reference: rqvae/models/rqtransformer/__init__.py:19-20"""
    return RQTransformer(config)
