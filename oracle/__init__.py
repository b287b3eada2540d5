"""CPU restatement of the VeritasFi retrieval hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package; nothing under veritasfi_b200/ does.  PARITY UNPINNED: the reference ships no golden
vectors for this path and its arithmetic lives in un-vendored libraries (see vfi_oracle.c header).
"""
from . import bm25, flat_ip, fusion, sharded  # noqa: F401
from ._clib import build as build_clib  # noqa: F401
