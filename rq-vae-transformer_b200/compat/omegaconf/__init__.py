"""Minimal `omegaconf` stand-in backed by rqvae.utils.config.Config (the real package is not installable offline).

Covers what the reference's sampling-path scripts call: OmegaConf.create / structured / merge /
from_cli / from_dotlist / to_yaml / to_container / save / load, MISSING, DictConfig.  Put
`rq-vae-transformer_b200/compat` on PYTHONPATH *after* the real omegaconf if you have it -- this is only a fallback."""
import dataclasses
import sys

import yaml

from rqvae.utils.config import MISSING, Config, merge as _merge

DictConfig = Config


def _parse(v):
    try:
        return yaml.safe_load(v)
    except Exception:
        return v


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if isinstance(obj, str):
            obj = yaml.safe_load(obj)
        return Config(obj or {})

    @staticmethod
    def structured(obj):
        if dataclasses.is_dataclass(obj):
            inst = obj() if isinstance(obj, type) else obj
            return Config(dataclasses.asdict(inst))
        return Config(obj)

    @staticmethod
    def merge(*cfgs):
        return _merge(*cfgs)

    @staticmethod
    def from_dotlist(items):
        out = Config()
        for it in items:
            if "=" not in it:
                continue
            key, val = it.split("=", 1)
            node = out
            parts = key.lstrip("-").split(".")
            for p in parts[:-1]:
                if p not in node:
                    node[p] = Config()
                node = node[p]
            node[parts[-1]] = _parse(val)
        return out

    @staticmethod
    def from_cli(args_list=None):
        return OmegaConf.from_dotlist(sys.argv[1:] if args_list is None else args_list)

    @staticmethod
    def to_container(cfg, resolve=True):
        return cfg.to_dict() if isinstance(cfg, Config) else cfg

    @staticmethod
    def to_yaml(cfg):
        return yaml.safe_dump(OmegaConf.to_container(cfg), sort_keys=False)

    @staticmethod
    def save(cfg, f):
        text = OmegaConf.to_yaml(cfg)
        if hasattr(f, "write"):
            f.write(text)
        else:
            with open(f, "w") as fh:
                fh.write(text)

    @staticmethod
    def load(f):
        if hasattr(f, "read"):
            return Config(yaml.safe_load(f))
        with open(f) as fh:
            return Config(yaml.safe_load(fh))


__all__ = ["OmegaConf", "DictConfig", "MISSING"]
