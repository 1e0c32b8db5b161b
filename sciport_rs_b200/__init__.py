"""sciport_rs_b200 — B200-native batched special functions behind sciport-rs's `special` API.

Only what the hot path needs lives here: `csrc/` holds the CUDA
kernels and the C-ABI launcher layer (libsciport_b200.so), `special` mirrors
the reference's `special` module (/root/reference/src/special/{mod,kv,trig}.rs)
and widens it with the batched J/Y/I/K/H1/H2/Airy/gamma entry points.
"""
from . import _lib  # noqa: F401  (loads the shared library; raises if it is missing)
from . import special  # noqa: F401

__all__ = ["special"]
