"""Abstract model roles (reference: rqvae/models/interfaces.py:20,53,71)."""
from torch import nn


class Stage1Model(nn.Module):
    """image <-> code model: get_codes / decode_code / get_recon_imgs"""

    def get_codes(self, *args, **kwargs):
        raise NotImplementedError

    def decode_code(self, *args, **kwargs):
        raise NotImplementedError

    def get_recon_imgs(self, *args, **kwargs):
        raise NotImplementedError


class Stage2Model(nn.Module):
    """code prior; exposes block_size through get_block_size() (interfaces.py:71)"""

    def get_block_size(self):
        return self.block_size
