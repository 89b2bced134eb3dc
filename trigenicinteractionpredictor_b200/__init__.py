"""B200-native MMSBM EM for trigenic-interaction prediction (drop-in for the reference `Model`).

    from trigenicinteractionpredictor_b200 import Model

The numerics live in libtip.so (hand-written CUDA for sm_100a behind include/tip.h); this package
is the host-side mirror of the reference's class surface plus the torch.distributed plumbing.
"""
from .TrigenicInteractionPredictor import Model, main, train_sample  # noqa: F401

__all__ = ["Model", "main", "train_sample"]
__version__ = "0.1.0"
