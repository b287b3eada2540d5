"""Make the UNMODIFIED reference files run on the GPU kernels: register module shims for the
third-party libraries they import (`faiss`, `bm25s`, `Stemmer`) backed by this package.

    import veritasfi_b200.dropin as dropin
    dropin.install()                  # before `from src.utils.ensembleRetriever import EnsembleRetriever`

After install(), /root/reference/src/utils/faissRetriever.py:3 (`import faiss`) and
bm25Retriever.py:7-8 (`import bm25s`, `import Stemmer`) bind to faiss_compat / bm25_compat.
See INTEGRATION.md for the two-line alternative (change the imports in those files)."""
from __future__ import annotations

import sys
import types

from . import bm25_compat, faiss_compat
from . import stemmer as native_stemmer


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install(override_stemmer: bool = False) -> None:
    sys.modules["faiss"] = _module(
        "faiss", IndexFlatIP=faiss_compat.IndexFlatIP, normalize_L2=faiss_compat.normalize_L2,
        METRIC_INNER_PRODUCT=faiss_compat.METRIC_INNER_PRODUCT, _vfi_shim=True)
    sys.modules["bm25s"] = _module(
        "bm25s", BM25=bm25_compat.BM25, tokenize=bm25_compat.tokenize, Tokenized=bm25_compat.Tokenized, _vfi_shim=True)
    have_real = False
    if not override_stemmer:
        try:
            import Stemmer  # noqa: F401
            have_real = not getattr(sys.modules["Stemmer"], "_vfi_shim", False)
        except Exception:
            have_real = False
    if not have_real:
        sys.modules["Stemmer"] = _module("Stemmer", Stemmer=native_stemmer.Stemmer, algorithms=native_stemmer.algorithms,
                                         _vfi_shim=True)


def uninstall() -> None:
    for name in ("faiss", "bm25s", "Stemmer"):
        m = sys.modules.get(name)
        if m is not None and getattr(m, "_vfi_shim", False):
            del sys.modules[name]
