"""nlspn_eccv20_b200 -- B200-native (sm_100a) NLSPN propagation: one hot path, nothing else.

    from nlspn_eccv20_b200 import NLSPN            # drop-in module (north-star signature)
    import nlspn_eccv20_b200.dcn as DCN            # drop-in for the reference's native extension

All arithmetic runs in ``lib/libnlspn_b200.so`` (hand-written CUDA behind a C ABI, see
include/nlspn_b200.h).  There is no CPU or PyTorch fallback: a missing library raises.
"""
from .nlspn import NLSPN, NLSPNFunction, nlspn_propagate, GraphedNLSPN  # noqa: F401
from . import dcn, functional  # noqa: F401

__version__ = "0.1.0"
