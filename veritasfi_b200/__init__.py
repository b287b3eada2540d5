"""veritasfi_b200 — B200-native (sm_100a) implementation of VeritasFi's multi-path retrieval hot path.

Everything computes on the GPU through the C ABI in include/vfi.h (libvfi.so, built by
`python -m veritasfi_b200.build`).  There is no CPU fallback: importing the compute modules without the
built library, or calling them without a CUDA device, raises.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401
from .build import build_lib  # noqa: F401
