"""Drop-in for the slice of PyStemmer the reference uses: `Stemmer.Stemmer('english')` with `stemWord` /
`stemWords` (src/utils/bm25Retriever.py:14,47 hand it to `bm25s.tokenize`, which calls `stemWords` on the
vocabulary).  The stemming itself is the native Snowball-English routine behind `vfi_stem_english`
(csrc/text_host.h); there is no Python implementation on this path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_ALGORITHMS = ("english", "en", "porter2")


def algorithms() -> list[str]:
    return ["english"]


def stem_words_native(words) -> list[str]:
    """One call of vfi_stem_english over the whole list (UTF-8 in, UTF-8 out)."""
    words = list(words)
    if not words:
        return []
    enc = [w.encode("utf-8") for w in words]
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in enc], out=offsets[1:])
    blob = b"".join(enc)
    out = C.create_string_buffer(max(1, len(blob)))
    out_off = np.zeros(len(enc) + 1, dtype=np.int64)
    N.check(N.load().vfi_stem_english(blob, offsets.ctypes.data, len(enc), C.addressof(out), len(blob),
                                      out_off.ctypes.data))
    raw = out.raw
    return [raw[out_off[i]:out_off[i + 1]].decode("utf-8", "replace") for i in range(len(enc))]


def tokenize_ascii_native(text: str):
    """Token strings of bm25s' pattern for ASCII text through vfi_tokenize_ascii; None when the text is not ASCII
    (the caller then applies the Unicode-aware pattern)."""
    if not text.isascii():
        return None
    raw = text.encode("ascii")
    cap = len(raw) // 2 + 1
    starts = np.empty(cap, dtype=np.int64)
    lens = np.empty(cap, dtype=np.int64)
    n = C.c_int64(0)
    N.check(N.load().vfi_tokenize_ascii(raw, len(raw), starts.ctypes.data, lens.ctypes.data, cap, C.byref(n)))
    return [text[int(s):int(s + l)] for s, l in zip(starts[:n.value], lens[:n.value])]


class Stemmer:
    """`Stemmer.Stemmer(algorithm)` of PyStemmer for algorithm = 'english'."""

    def __init__(self, algorithm: str = "english", maxCacheSize: int = 10000):
        if algorithm.lower() not in _ALGORITHMS:
            raise KeyError(f"Stemming algorithm '{algorithm}' not found")   # PyStemmer raises KeyError too
        self.algorithm = "english"
        self.maxCacheSize = maxCacheSize

    def stemWord(self, word: str) -> str:
        return stem_words_native([word])[0]

    def stemWords(self, words) -> list[str]:
        return stem_words_native(words)
