"""Drop-in for the slice of the `bm25s` module the reference uses.

Reference call sites (paths relative to /root/reference/):
    bm25s.tokenize(corpus, stopwords="english", stemmer=stemmer)      src/utils/bm25Retriever.py:15,67
    retriever = bm25s.BM25(); retriever.index(tokens); .save(dir, corpus=doc_ids)   :16-18
    bm25s.BM25.load(dir_path, load_corpus=True); engine.corpus                      :46,48
    engine.retrieve(query_tokens, k=k, return_as="tuple") -> (docs, scores)         :75-79

Online scoring (`retrieve`) runs on the GPU through the C ABI (vfi_bm25_search / vfi_bm25_rank_all);
index construction and tokenisation are the offline/host steps of the reference (SURVEY.md §8a
row B1) and stay host-side numpy/regex like bm25s itself.  Statements about bm25s internals are
[upstream] knowledge — the library is not vendored in the reference nor installed here.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import re
from collections import namedtuple
from typing import Iterable, Sequence

import numpy as np

from . import _native as N

# [upstream] bm25s.tokenization STOPWORDS_EN
STOPWORDS_EN = (
    "a", "an", "and", "are", "as", "at", "be", "but", "by", "for", "if", "in", "into", "is", "it",
    "no", "not", "of", "on", "or", "such", "that", "the", "their", "then", "there", "these",
    "they", "this", "to", "was", "will", "with",
)
_TOKEN_RE = re.compile(r"(?u)\b\w\w+\b")

Tokenized = namedtuple("Tokenized", ["ids", "vocab"])
Results = namedtuple("Results", ["documents", "scores"])


def _stem(stemmer, words: list[str]) -> list[str]:
    if stemmer is None:
        return words
    if hasattr(stemmer, "stemWords"):
        return list(stemmer.stemWords(words))
    if callable(stemmer):
        return list(stemmer(words))
    raise TypeError("stemmer must expose stemWords(list) or be callable")


def _split(text: str) -> list[str]:
    """Token strings of the bm25s pattern: the native ASCII splitter (vfi_tokenize_ascii) when the text is ASCII,
    the Unicode-aware regular expression otherwise (both are host-side and give the same tokens on ASCII text)."""
    if text.isascii():
        from .stemmer import tokenize_ascii_native
        toks = tokenize_ascii_native(text)
        if toks is not None:
            return toks
    return _TOKEN_RE.findall(text)


def tokenize(texts, stopwords="english", stemmer=None, lower: bool = True, return_ids: bool = True,
             show_progress: bool = False, **_ignored):
    """bm25s.tokenize: lower-case, regex `\\b\\w\\w+\\b`, stop-word removal, optional stemming of the
    unique tokens, ids assigned in first-seen order."""
    if isinstance(texts, str):
        texts = [texts]
    if stopwords in ("english", "en", True):
        stop = set(STOPWORDS_EN)
    elif not stopwords:
        stop = set()
    else:
        stop = set(stopwords)
    vocab: dict[str, int] = {}
    ids: list[list[int]] = []
    for text in texts:
        if lower:
            text = text.lower()
        doc = []
        for tok in _split(text):
            if tok in stop:
                continue
            if tok not in vocab:
                vocab[tok] = len(vocab)
            doc.append(vocab[tok])
        ids.append(doc)
    if stemmer is not None:
        words = list(vocab)
        stemmed = _stem(stemmer, words)
        new_vocab: dict[str, int] = {}
        remap = np.empty(len(words), dtype=np.int64)
        for w, s in zip(words, stemmed):
            if s not in new_vocab:
                new_vocab[s] = len(new_vocab)
            remap[vocab[w]] = new_vocab[s]
        ids = [[int(remap[i]) for i in doc] for doc in ids]
        vocab = new_vocab
    if not return_ids:
        rev = {i: w for w, i in vocab.items()}
        return [[rev[i] for i in doc] for doc in ids]
    return Tokenized(ids=ids, vocab=vocab)


def build_csc(doc_ptr: np.ndarray, doc_tokens: np.ndarray, n_vocab: int, k1: float = 1.5, b: float = 0.75,
              method: str = "lucene"):
    """Token-major posting arrays from a CSR of token ids per document — bm25s.BM25.index [upstream].

    impact(t, d) = idf(t) * tf / (tf + k1 * (1 - b + b * dl / avgdl)),  idf = ln(1 + (N - df + 0.5)/(df + 0.5)),
    computed in float64 and stored as float32; postings of a token are sorted by doc id.
    Returns (indptr int64 [V+1], indices int32 [nnz], data float32 [nnz])."""
    if method != "lucene":
        raise NotImplementedError("only the default bm25s method 'lucene' is restated")
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    doc_tokens = np.asarray(doc_tokens, dtype=np.int64)
    n_docs = len(doc_ptr) - 1
    dl = np.diff(doc_ptr).astype(np.float64)
    avgdl = float(dl.mean()) if n_docs else 0.0
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), np.diff(doc_ptr))
    pair = doc_tokens * n_docs + doc_of                       # token-major, doc ascending
    uniq, tf = np.unique(pair, return_counts=True)
    tok = uniq // max(n_docs, 1)
    doc = uniq - tok * max(n_docs, 1)
    df = np.bincount(tok, minlength=n_vocab).astype(np.float64)
    idf = np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5))
    tf = tf.astype(np.float64)
    tfc = tf / (tf + k1 * (1.0 - b + b * dl[doc] / (avgdl if avgdl > 0 else 1.0)))
    data = (idf[tok] * tfc).astype(np.float32)
    indptr = np.zeros(n_vocab + 1, dtype=np.int64)
    np.cumsum(np.bincount(tok, minlength=n_vocab), out=indptr[1:])
    return indptr, doc.astype(np.int32), data


class GpuPostings:
    """Token-major postings of one doc shard resident in HBM (handle on vfi_bm25_t)."""

    def __init__(self, indptr, indices, data, n_docs: int, id_offset: int = 0, device: int = 0):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        self.n_vocab = len(self.indptr) - 1
        self.n_docs = int(n_docs)
        self.id_offset = int(id_offset)
        self._h = C.c_void_p()
        N.check(N.load().vfi_bm25_create(self.indptr.ctypes.data_as(C.c_void_p), self.indices.ctypes.data_as(C.c_void_p),
                                         self.data.ctypes.data_as(C.c_void_p), self.n_vocab, self.n_docs,
                                         self.id_offset, int(device), C.byref(self._h)))

    def search(self, token_lists: Sequence[Sequence[int]], k: int):
        """token_lists: per query, vocabulary ids in query order (unknown tokens already dropped).
        Returns (ids int64 [B,k], scores float32 [B,k]) host arrays."""
        B = len(token_lists)
        qptr = np.zeros(B + 1, dtype=np.int64)
        np.cumsum([len(t) for t in token_lists], out=qptr[1:])
        toks = np.fromiter((t for lst in token_lists for t in lst), dtype=np.int32, count=int(qptr[-1]))
        return self.search_csr(toks, qptr, k)

    def search_csr(self, toks: np.ndarray, qptr: np.ndarray, k: int):
        B = len(qptr) - 1
        toks = np.ascontiguousarray(toks, dtype=np.int32)
        qptr = np.ascontiguousarray(qptr, dtype=np.int64)
        scores = np.full((B, k), -np.finfo(np.float32).max, dtype=np.float32)
        ids = np.full((B, k), -1, dtype=np.int64)
        if B:
            N.check(N.load().vfi_bm25_search(self._h, toks.ctypes.data_as(C.c_void_p), qptr.ctypes.data_as(C.c_void_p),
                                             B, int(k), scores.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p),
                                             N.MEM_HOST, None))
        return ids, scores

    def search_csr_device(self, toks: np.ndarray, qptr: np.ndarray, k: int, device=None):
        """Same search with the results left on the GPU: (ids int64 [B,k], scores float32 [B,k]) torch CUDA tensors
        (the hybrid path hands them straight to the fusion kernel)."""
        import torch

        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        B = len(qptr) - 1
        toks = np.ascontiguousarray(toks, dtype=np.int32)
        qptr = np.ascontiguousarray(qptr, dtype=np.int64)
        scores = torch.full((B, k), -float(np.finfo(np.float32).max), dtype=torch.float32, device=dev)
        ids = torch.full((B, k), -1, dtype=torch.int64, device=dev)
        if B:
            N.check(N.load().vfi_bm25_search(self._h, toks.ctypes.data_as(C.c_void_p), qptr.ctypes.data_as(C.c_void_p),
                                             B, int(k), C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()),
                                             N.MEM_DEVICE, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return ids, scores

    @staticmethod
    def pack_tokens(token_lists: Sequence[Sequence[int]]):
        """Per-query token id lists -> (toks int32 [total], qptr int64 [B+1]) as vfi_bm25_search takes them."""
        qptr = np.zeros(len(token_lists) + 1, dtype=np.int64)
        np.cumsum([len(t) for t in token_lists], out=qptr[1:])
        toks = np.fromiter((t for lst in token_lists for t in lst), dtype=np.int32, count=int(qptr[-1]))
        return toks, qptr

    def score_all(self, tokens: Sequence[int]) -> np.ndarray:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.zeros(self.n_docs, dtype=np.float32)
        N.check(N.load().vfi_bm25_score_all(self._h, toks.ctypes.data_as(C.c_void_p), len(toks),
                                            out.ctypes.data_as(C.c_void_p), N.MEM_HOST, None))
        return out

    def rank_all(self, tokens: Sequence[int]):
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        scores = np.zeros(self.n_docs, dtype=np.float32)
        ids = np.zeros(self.n_docs, dtype=np.int64)
        N.check(N.load().vfi_bm25_rank_all(self._h, toks.ctypes.data_as(C.c_void_p), len(toks),
                                           scores.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p), None))
        return ids, scores

    def set_profile(self, on: bool) -> None:
        N.check(N.load().vfi_bm25_set_profile(self._h, int(on)))

    def stats(self, reset: bool = False) -> N.Bm25Stats:
        st = N.Bm25Stats()
        N.check(N.load().vfi_bm25_get_stats(self._h, C.byref(st), int(reset)))
        return st

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.load().vfi_bm25_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BM25:
    """bm25s.BM25 façade: same constructor defaults (k1=1.5, b=0.75, method="lucene"), index / save /
    load / retrieve; `retrieve` scores on the GPU."""

    def __init__(self, k1: float = 1.5, b: float = 0.75, delta: float = 0.5, method: str = "lucene",
                 idf_method=None, dtype: str = "float32", int_dtype: str = "int32", corpus=None,
                 backend: str = "cuda", device: int = 0):
        self.k1, self.b, self.delta, self.method = k1, b, delta, method
        self.dtype, self.int_dtype = dtype, int_dtype
        self.corpus = corpus
        self.device = device
        self.vocab_dict: dict[str, int] = {}
        self.scores: dict | None = None
        self._gpu: GpuPostings | None = None

    # ---- offline ----------------------------------------------------------------------------
    def index(self, corpus, show_progress: bool = False, **_ignored) -> None:
        """corpus: Tokenized(ids, vocab) as returned by tokenize(), or a list of token-string lists."""
        if isinstance(corpus, Tokenized) or (isinstance(corpus, tuple) and len(corpus) == 2):
            ids, vocab = corpus
        else:
            vocab = {}
            ids = []
            for doc in corpus:
                row = []
                for tok in doc:
                    if tok not in vocab:
                        vocab[tok] = len(vocab)
                    row.append(vocab[tok])
                ids.append(row)
        doc_ptr = np.zeros(len(ids) + 1, dtype=np.int64)
        np.cumsum([len(d) for d in ids], out=doc_ptr[1:])
        toks = np.fromiter((t for d in ids for t in d), dtype=np.int64, count=int(doc_ptr[-1]))
        indptr, indices, data = build_csc(doc_ptr, toks, len(vocab), self.k1, self.b, self.method)
        self.vocab_dict = dict(vocab)
        self.scores = {"data": data, "indices": indices, "indptr": indptr, "num_docs": len(ids)}
        self._gpu = None

    def save(self, save_dir, corpus=None, **_ignored) -> None:
        """[upstream] bm25s layout: {data,indices,indptr}.csc.index.npy, vocab.index.json,
        params.index.json, corpus.jsonl (+ corpus.mmindex.json line offsets)."""
        if self.scores is None:
            raise ValueError("index() must be called before save()")
        os.makedirs(save_dir, exist_ok=True)
        np.save(os.path.join(save_dir, "data.csc.index.npy"), self.scores["data"], allow_pickle=False)
        np.save(os.path.join(save_dir, "indices.csc.index.npy"), self.scores["indices"].astype(np.int32), allow_pickle=False)
        np.save(os.path.join(save_dir, "indptr.csc.index.npy"), self.scores["indptr"].astype(np.int32), allow_pickle=False)
        with open(os.path.join(save_dir, "vocab.index.json"), "w", encoding="utf-8") as f:
            json.dump(self.vocab_dict, f)
        params = dict(k1=self.k1, b=self.b, delta=self.delta, method=self.method, idf_method=self.method,
                      dtype=self.dtype, int_dtype=self.int_dtype, num_docs=self.scores["num_docs"], version="vfi-b200",
                      backend="cuda")
        with open(os.path.join(save_dir, "params.index.json"), "w", encoding="utf-8") as f:
            json.dump(params, f, indent=4)
        corpus = corpus if corpus is not None else self.corpus
        if corpus is not None:
            offsets = []
            with open(os.path.join(save_dir, "corpus.jsonl"), "w", encoding="utf-8") as f:
                for i, doc in enumerate(corpus):
                    offsets.append(f.tell())
                    if isinstance(doc, str):
                        doc = {"id": i, "text": doc}
                    elif isinstance(doc, dict):
                        doc = dict(doc)
                        doc.setdefault("id", i)
                    else:
                        doc = {"id": i, "text": doc}
                    f.write(json.dumps(doc, ensure_ascii=False) + "\n")
            with open(os.path.join(save_dir, "corpus.mmindex.json"), "w", encoding="utf-8") as f:
                json.dump(offsets, f)

    @classmethod
    def load(cls, save_dir, load_corpus: bool = False, mmap: bool = False, device: int = 0, **_ignored) -> "BM25":
        with open(os.path.join(save_dir, "params.index.json"), encoding="utf-8") as f:
            params = json.load(f)
        self = cls(k1=params.get("k1", 1.5), b=params.get("b", 0.75), delta=params.get("delta", 0.5),
                   method=params.get("method", "lucene"), device=device)
        with open(os.path.join(save_dir, "vocab.index.json"), encoding="utf-8") as f:
            self.vocab_dict = json.load(f)
        mode = "r" if mmap else None
        data = np.load(os.path.join(save_dir, "data.csc.index.npy"), mmap_mode=mode)
        indices = np.load(os.path.join(save_dir, "indices.csc.index.npy"), mmap_mode=mode)
        indptr = np.load(os.path.join(save_dir, "indptr.csc.index.npy"), mmap_mode=mode)
        self.scores = {"data": np.asarray(data, dtype=np.float32), "indices": np.asarray(indices, dtype=np.int32),
                       "indptr": np.asarray(indptr, dtype=np.int64), "num_docs": int(params["num_docs"])}
        if load_corpus:
            path = os.path.join(save_dir, "corpus.jsonl")
            corpus = []
            if os.path.exists(path):
                with open(path, encoding="utf-8") as f:
                    for line in f:
                        line = line.strip()
                        if line:
                            corpus.append(json.loads(line))
            self.corpus = corpus
        return self

    # ---- online -----------------------------------------------------------------------------
    def _postings(self) -> GpuPostings:
        if self.scores is None:
            raise ValueError("the index is empty: call index() or load() first")
        if self._gpu is None:
            s = self.scores
            self._gpu = GpuPostings(s["indptr"], s["indices"], s["data"], s["num_docs"], 0, self.device)
        return self._gpu

    def get_tokens_ids(self, query_tokens: Iterable[str]) -> list[int]:
        return [self.vocab_dict[t] for t in query_tokens if t in self.vocab_dict]

    def _query_id_lists(self, query_tokens) -> list[list[int]]:
        if isinstance(query_tokens, Tokenized) or (isinstance(query_tokens, tuple) and len(query_tokens) == 2
                                                   and isinstance(query_tokens[1], dict)):
            ids, vocab = query_tokens
            rev = {i: w for w, i in vocab.items()}
            return [self.get_tokens_ids(rev[i] for i in q) for q in ids]
        out = []
        for q in query_tokens:
            if len(q) and isinstance(q[0], str):
                out.append(self.get_tokens_ids(q))
            else:
                out.append([int(t) for t in q if 0 <= int(t) < len(self.scores["indptr"]) - 1])
        return out

    def retrieve(self, query_tokens, corpus=None, k: int = 10, sorted: bool = True, return_as: str = "tuple",
                 show_progress: bool = False, **_ignored):
        """Top-k per query.  Returns Results(documents [nq,k], scores float32 [nq,k]); documents are corpus
        entries when a corpus is attached, else row ids.  Raises ValueError if k > number of docs
        (bm25s behaviour [upstream])."""
        gp = self._postings()
        n_docs = gp.n_docs
        if k > n_docs:
            raise ValueError(f"k of {k} is larger than the number of available scores, which is {n_docs} "
                             "(corpus size should be larger than top-k). Please set with a smaller k or increase the size of corpus.")
        lists = self._query_id_lists(query_tokens)
        nq = len(lists)
        if k <= N.MAX_K // 2:
            ids, scores = gp.search(lists, k)
        else:  # k up to N: rank every doc (ensembleRetriever.py:189 asks for k = num_chunk)
            ids = np.empty((nq, k), dtype=np.int64)
            scores = np.empty((nq, k), dtype=np.float32)
            for i, toks in enumerate(lists):
                ri, rs = gp.rank_all(toks)
                ids[i], scores[i] = ri[:k], rs[:k]
        corpus = corpus if corpus is not None else self.corpus
        if corpus is not None:
            docs = np.empty((nq, k), dtype=object)
            for i in range(nq):
                for j in range(k):
                    docs[i, j] = corpus[int(ids[i, j])]
        else:
            docs = ids
        if return_as == "tuple":
            return Results(documents=docs, scores=scores)
        if return_as == "documents":
            return docs
        raise ValueError("return_as must be 'tuple' or 'documents'")
