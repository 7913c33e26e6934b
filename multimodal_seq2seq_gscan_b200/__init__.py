"""B200-native implementation of the gSCAN multimodal seq2seq training / greedy-decoding hot path.

``Model`` is a drop-in for the reference's ``seq2seq.model.Model``; its arithmetic runs in
``lib/libgscan_b200.so`` (hand-written sm_100a CUDA behind the C ABI in ``include/gscan_b200.h``).
"""
from .model import Model  # noqa: F401
from . import ops  # noqa: F401
from ._lib import build, load, LIB_PATH  # noqa: F401

__all__ = ["Model", "ops", "build", "load", "LIB_PATH"]
