"""create_model factory (reference: rqvae/models/__init__.py:20-37)."""
from .rqvae import get_rqvae
from .rqtransformer import get_rqtransformer


def create_model(config, ema=False):
    kind = str(config.type).lower()
    builders = {"rq-transformer": get_rqtransformer, "rq-vae": get_rqvae}
    if kind not in builders:
        raise ValueError(f"{kind} is invalid..")
    model = builders[kind](config)
    model_ema = None
    if ema:
        # sampling never updates an EMA; the EMA copy is just a second instance that receives `state_dict_ema`
        # (main_sampling_fid.py:153-155).  Training-time EMA tracking is out of scope.
        from .ema import ExponentialMovingAverage
        model_ema = ExponentialMovingAverage(builders[kind](config), getattr(config, "ema", 0.9999))
        model_ema.eval()
        model_ema.update(model, step=-1)
    return model, model_ema
