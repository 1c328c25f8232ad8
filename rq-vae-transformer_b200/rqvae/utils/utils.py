"""Sampler + seeding helpers (reference: rqvae/utils/utils.py:41-48, 60-123)."""
import pickle
import random

import numpy as np
import torch

from ..models import _bind as nb


def set_seed(seed=None):
    """utils.py:41-48"""
    if seed is None:
        seed = random.getrandbits(32)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    return seed


def save_pickle(fname, data):
    with open(fname, "wb") as fp:
        pickle.dump(data, fp, protocol=pickle.HIGHEST_PROTOCOL)


def load_pickle(fname):
    with open(fname, "rb") as fp:
        return pickle.load(fp)


@torch.no_grad()
def sample_from_logits(logits, temperature=1.0, top_k=None, top_p=None):
    """utils.py:82-123 -- one fused kernel (csrc/sampler.cu).  The Exp(1) noise is drawn here with one
    ``exponential_`` call of shape [B,V] from the device's default generator -- exactly the draw
    ``torch.multinomial(probs, 1)`` makes -- so results are RNG-stream identical to the reference."""
    q = torch.empty(logits.shape, dtype=torch.float32, device=logits.device).exponential_(1)
    return nb.sample_logits(logits, temperature, top_k, top_p, q=q)


def top_k_logits(logits, k):
    raise NotImplementedError("rqb200: fused into sample_from_logits (csrc/sampler.cu)")


def top_p_probs(probs, p):
    raise NotImplementedError("rqb200: fused into sample_from_logits (csrc/sampler.cu)")
