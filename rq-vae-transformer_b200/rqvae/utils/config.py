"""Config layer: yaml -> attribute dict with deep merge (reference: rqvae/utils/config.py:17-49).

The reference goes yaml -> EasyDict -> OmegaConf; neither package is installable here, so ``Config`` provides the
subset the sampling path and ``measure_throughput`` use: attribute + item access, nested assignment, deep ``copy()``,
``**`` unpacking, ``merge``.  ``rq-vae-transformer_b200/compat/`` additionally ships ``omegaconf`` / ``easydict`` shim
modules backed by this class so the reference's unchanged scripts import."""
import copy

import yaml

MISSING = "???"


class Config(dict):
    def __init__(self, data=None, **kw):
        super().__init__()
        self.update_from(data or {})
        self.update_from(kw)

    @staticmethod
    def _wrap(v):
        if isinstance(v, Config):
            return v
        if isinstance(v, dict):
            return Config(v)
        if isinstance(v, (list, tuple)):
            return [Config._wrap(x) for x in v]
        return v

    def update_from(self, other):
        """deep merge: nested dicts merge key-wise, everything else overwrites"""
        for k, v in dict(other).items():
            if isinstance(v, dict) and isinstance(self.get(k), Config):
                self[k].update_from(v)
            else:
                self[k] = Config._wrap(copy.deepcopy(v) if isinstance(v, (dict, list)) else v)
        return self

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = Config._wrap(v)

    def __setitem__(self, k, v):
        super().__setitem__(k, Config._wrap(v))

    def copy(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def get(self, k, default=None):
        return self[k] if k in self else default

    def to_dict(self):
        def un(v):
            if isinstance(v, Config):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, list):
                return [un(x) for x in v]
            return v
        return un(self)


def merge(*configs):
    out = Config()
    for c in configs:
        out.update_from(c)
    return out


def load_config(config_path):
    """config.py:17-22"""
    with open(config_path) as f:
        return Config(yaml.load(f, Loader=yaml.FullLoader))


def is_stage1_arch(arch_type):
    return "transformer" not in arch_type


def augment_arch_defaults(arch_config):
    """config.py:29-49"""
    if arch_config.type == "rq-vae":
        defaults = Config({"ema": None,
                           "hparams": {"loss_type": "l1", "restart_unused_codes": False, "use_padding_idx": False,
                                       "masked_dropout": 0.0},
                           "checkpointing": False})
        return merge(defaults, arch_config)
    if arch_config.type == "rq-transformer":
        from ..models.rqtransformer.configs import RQTransformerConfig
        return RQTransformerConfig.create(arch_config)
    raise NotImplementedError
