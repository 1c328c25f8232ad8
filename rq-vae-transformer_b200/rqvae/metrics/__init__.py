"""Import-compatible placeholders for rqvae.metrics (reference: rqvae/metrics/__init__.py:15-17).

FID / IS / CLIP-score need pretrained Inception / CLIP weights and dataset statistics that are not available offline and are
outside the hot path.  The sampling scripts import these names at module load; they only call them
when statistics are requested (`--no-stats-saving` skips them, main_sampling_fid.py:256)."""
from .fid import compute_fid, compute_rfid, compute_statistics_from_files


def _unavailable(name):
    def fn(*a, **k):
        raise NotImplementedError("rqb200: %s is out of scope (needs pretrained networks / dataset statistics); run the "
                                  "sampling script with --no-stats-saving" % name)
    fn.__name__ = name
    return fn


compute_IS = _unavailable("compute_IS")
compute_clip_score = _unavailable("compute_clip_score")
