"""gstreamer_vit_tracker_b200 — B200-native per-frame hot path of a GStreamer ViT tracker.

Hand-written sm_100a CUDA behind a C ABI (include/vt_tracker.h, libvittrack_b200.so) plus a thin
Python mirror of the reference's interface (api.py).  There is no CPU / PyTorch fallback: if the
CUDA library is not built, importing the API raises.
"""
from . import synth, weights  # noqa: F401  (pure-numpy input generation / weight files)

__all__ = ["synth", "weights", "api"]


def __getattr__(name):
    if name == "api":
        import importlib

        return importlib.import_module(".api", __name__)
    raise AttributeError(name)
