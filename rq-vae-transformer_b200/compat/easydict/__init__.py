"""Minimal `easydict` stand-in (attribute dict) backed by rqvae.utils.config.Config."""
from rqvae.utils.config import Config as EasyDict

__all__ = ["EasyDict"]
