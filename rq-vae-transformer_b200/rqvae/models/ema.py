"""Minimal EMA wrapper so checkpoints with `state_dict_ema` load through the same surface
(reference: rqvae/models/ema.py:21-54).  Only what sampling needs: `.module`, attribute forwarding, `update`."""
import torch
from torch import nn


class ExponentialMovingAverage(nn.Module):
    def __init__(self, init_module, mu):
        super().__init__()
        self.module = init_module
        self.mu = mu

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    @torch.no_grad()
    def update(self, module, step=None):
        # rqvae/models/ema.py:30-43: step None -> self.mu; step < 0 -> hard copy; else the warm-up schedule
        mu = self.mu if step is None else (0.0 if step < 0 else min(self.mu, (1.0 + step) / (10.0 + step)))
        src = dict(module.state_dict())
        for name, dst in self.module.state_dict().items():
            s = src[name].to(dst.device)
            if dst.dtype.is_floating_point:
                dst.mul_(mu).add_(s, alpha=1.0 - mu)
            else:
                dst.copy_(s)
        if hasattr(self.module, "_invalidate_native"):
            self.module._invalidate_native()

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)
