"""Oracle for the sparse path: bm25s index construction and retrieve as the reference calls them
(/root/reference/src/utils/bm25Retriever.py:14-18,67,75-79) [upstream semantics restated from the
published algorithm].  TEST INFRASTRUCTURE ONLY — see oracle/vfi_oracle.c."""
from __future__ import annotations

import numpy as np

from ._clib import lib, ptr


def build_index(docs, n_vocab: int, k1: float = 1.5, b: float = 0.75):
    """docs: list of token-id lists.  Lucene-variant BM25 impacts precomputed per (token, doc) and
    stored token-major with doc ids ascending (bm25s.BM25.index, method="lucene"):
        idf = ln(1 + (N - df + 0.5)/(df + 0.5));  impact = idf * tf / (tf + k1*(1 - b + b*dl/avgdl))
    Written as plain loops on purpose (independent of the vectorised product builder)."""
    n_docs = len(docs)
    dl = [len(d) for d in docs]
    avgdl = (sum(dl) / n_docs) if n_docs else 0.0
    postings: list[dict[int, int]] = [dict() for _ in range(n_vocab)]
    for di, doc in enumerate(docs):
        for t in doc:
            postings[t][di] = postings[t].get(di, 0) + 1
    indptr = np.zeros(n_vocab + 1, dtype=np.int64)
    indices, data = [], []
    for t in range(n_vocab):
        df = len(postings[t])
        idf = np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5))
        for di in sorted(postings[t]):
            tf = float(postings[t][di])
            tfc = tf / (tf + k1 * (1.0 - b + b * dl[di] / (avgdl if avgdl > 0 else 1.0)))
            indices.append(di)
            data.append(np.float32(idf * tfc))
        indptr[t + 1] = len(indices)
    return indptr, np.asarray(indices, dtype=np.int32), np.asarray(data, dtype=np.float32)


def scores(indptr, indices, data, tokens, n_docs: int) -> np.ndarray:
    """All doc scores of one query: zeros(N) then np.add.at per query token in order (fp32)."""
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float32)
    tokens = np.ascontiguousarray(tokens, dtype=np.int32)
    out = np.zeros(n_docs, dtype=np.float32)
    lib().vfo_bm25_scores(ptr(indptr), ptr(indices), ptr(data), len(indptr) - 1, ptr(tokens), len(tokens), n_docs, ptr(out))
    return out


def scores_numpy(indptr, indices, data, tokens, n_docs: int) -> np.ndarray:
    """The same thing literally as bm25s' numpy backend writes it (np.add.at) — cross-check and the
    timed CPU baseline of bench.py."""
    out = np.zeros(n_docs, dtype=np.float32)
    nv = len(indptr) - 1
    for t in tokens:
        if 0 <= t < nv:
            s, e = indptr[t], indptr[t + 1]
            np.add.at(out, indices[s:e], data[s:e])
    return out


def retrieve(indptr, indices, data, token_lists, n_docs: int, k: int, id_base: int = 0):
    """Top-k per query under (score desc, id asc).  Returns (ids [B,k], scores [B,k])."""
    from .flat_ip import topk

    B = len(token_lists)
    I = np.empty((B, k), np.int64)
    S = np.empty((B, k), np.float32)
    for q, toks in enumerate(token_lists):
        s = scores(indptr, indices, data, toks, n_docs)
        S[q], I[q] = topk(s, k, id_base)
    return I, S


# ---- tokenisation and the on-disk layout, restated independently of the product (bm25s [upstream]) -------------------
# bm25s.tokenization.STOPWORDS_EN [upstream]
STOPWORDS_EN = frozenset(
    "a an and are as at be but by for if in into is it no not of on or such that the their then there these they this to was "
    "will with".split())


def tokenize(texts, stemmer=None):
    """bm25s.tokenize(texts, stopwords="english", stemmer=stemmer) as bm25Retriever.py:15,67 calls it: lower-case, the
    pattern r"(?u)\\b\\w\\w+\\b", English stop words dropped, ids in first-seen order; with a stemmer the UNIQUE tokens are
    stemmed and ids re-assigned in first-seen order of the stems.  Returns (ids per text, vocab dict)."""
    import re
    pat = re.compile(r"(?u)\b\w\w+\b")
    vocab, out = {}, []
    for text in texts:
        row = []
        for tok in pat.findall(text.lower()):
            if tok in STOPWORDS_EN:
                continue
            if tok not in vocab:
                vocab[tok] = len(vocab)
            row.append(vocab[tok])
        out.append(row)
    if stemmer is not None:
        words = sorted(vocab, key=vocab.get)
        stems = stemmer.stemWords(words)
        new_vocab, remap = {}, {}
        for w, s in zip(words, stems):
            if s not in new_vocab:
                new_vocab[s] = len(new_vocab)
            remap[vocab[w]] = new_vocab[s]
        out = [[remap[i] for i in row] for row in out]
        vocab = new_vocab
    return out, vocab


def save_dir(path, indptr, indices, data, vocab, n_docs, corpus):
    """The bm25s index directory [upstream]: {data,indices,indptr}.csc.index.npy, vocab.index.json, params.index.json,
    corpus.jsonl (+ corpus.mmindex.json)."""
    import json
    import os
    os.makedirs(path, exist_ok=True)
    np.save(os.path.join(path, "data.csc.index.npy"), np.asarray(data, np.float32), allow_pickle=False)
    np.save(os.path.join(path, "indices.csc.index.npy"), np.asarray(indices, np.int32), allow_pickle=False)
    np.save(os.path.join(path, "indptr.csc.index.npy"), np.asarray(indptr, np.int32), allow_pickle=False)
    with open(os.path.join(path, "vocab.index.json"), "w", encoding="utf-8") as f:
        json.dump(vocab, f)
    with open(os.path.join(path, "params.index.json"), "w", encoding="utf-8") as f:
        json.dump(dict(k1=1.5, b=0.75, delta=0.5, method="lucene", idf_method="lucene", dtype="float32", int_dtype="int32",
                       num_docs=int(n_docs), version="oracle", backend="numpy"), f, indent=4)
    offsets = []
    with open(os.path.join(path, "corpus.jsonl"), "w", encoding="utf-8") as f:
        for i, doc in enumerate(corpus):
            offsets.append(f.tell())
            f.write(json.dumps({"id": i, "text": doc}, ensure_ascii=False) + "\n")
    with open(os.path.join(path, "corpus.mmindex.json"), "w", encoding="utf-8") as f:
        json.dump(offsets, f)


def load_dir(path):
    import json
    import os
    with open(os.path.join(path, "params.index.json"), encoding="utf-8") as f:
        params = json.load(f)
    with open(os.path.join(path, "vocab.index.json"), encoding="utf-8") as f:
        vocab = json.load(f)
    corpus = []
    with open(os.path.join(path, "corpus.jsonl"), encoding="utf-8") as f:
        for line in f:
            if line.strip():
                corpus.append(json.loads(line))
    return dict(indptr=np.load(os.path.join(path, "indptr.csc.index.npy")).astype(np.int64),
                indices=np.load(os.path.join(path, "indices.csc.index.npy")).astype(np.int32),
                data=np.load(os.path.join(path, "data.csc.index.npy")).astype(np.float32),
                vocab=vocab, n_docs=int(params["num_docs"]), corpus=corpus)
